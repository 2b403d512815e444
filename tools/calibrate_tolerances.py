"""Calibration of tests/tolerances.json ("to be calibrated on first run and then frozen").

Two steps, both committed so the frozen numbers can be re-derived:

  1. CPU (no GPU needed):   python tools/calibrate_tolerances.py eref
     e_ref per layer shape of tests/test_gpu_parity.py::SHAPES = max |oracle_fp32 - oracle_fp64| of the direct-form oracle
     (oracle/bipartite.py) on the test's own seeded inputs -- the error the "reference Python path" itself would show in
     fp32.  Stored under "e_ref"; the per-layer bound is max(4 e_ref, atol + rtol |y64|).

  2. GPU:  GF_PARITY_LOG=gpurun_out/parity_log.jsonl python -m pytest tests -m gpu -q
           python tools/calibrate_tolerances.py measured gpurun_out/parity_log.jsonl
     Every comparison of the suite appends one JSON line (what, kernel path, max-abs error, ratio to the applied bound, ratio
     to the CONTRACT bound 1e-4 + 2e-3 |y64|, relative RMS).  This step summarises them per kernel path into "measured"
     (worst contract ratio, worst relative RMS, worst end-to-end PSNR ...).  The frozen "layer" / "e2e" numbers are then
     set by hand to the measured worst case times a safety margin (>= 1.5x) and justified in DESIGN.md section 5.

oracle = in-repo restatement; reference source unavailable; parity unpinned.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TOL = os.path.join(ROOT, "tests", "tolerances.json")


def shape_key(s):
    return "C%d-%dx%d-k%d-%s-%s" % (s[0], s[1], s[2], s[3], s[6], s[7])


def eref():
    import torch
    from oracle import bipartite as ob
    from tests.test_gpu_parity import SHAPES
    out = {}
    for duplex in (False, True):
        for shape in SHAPES:
            C, H, W, k, D, p, integration, norm = shape
            g = torch.Generator().manual_seed(C + H + k + (1 if duplex else 0))
            x = torch.randn(2, C, H, W, generator=g, dtype=torch.float64) * 1.3 + 0.2
            y = torch.randn(2, k, D, generator=g, dtype=torch.float64)
            w = ob.init_params(C, D, k, p, integration, duplex, seed=8 if duplex else 7, bias_std=0.4)
            nrm = None if norm == "none" else norm
            r64, _, _ = ob.transformer_layer(x, y, w, integration=integration, norm=nrm, duplex=duplex)
            w32 = {n: t.float() for n, t in w.items()}
            r32, _, _ = ob.transformer_layer(x.float(), y.float(), w32, integration=integration, norm=nrm, duplex=duplex)
            e = (r32.double() - r64).abs().max().item()
            out[("duplex/" if duplex else "simplex/") + shape_key(shape)] = e
            print(f"{'duplex ' if duplex else 'simplex'} {shape_key(shape):40s} e_ref = {e:.3e}  (|y|max {r64.abs().max().item():.2f})")
    tol = json.load(open(TOL))
    tol["e_ref"] = out
    json.dump(tol, open(TOL, "w"), indent=2)


def measured(log):
    recs = [json.loads(l) for l in open(log) if l.strip()]
    summ = {}
    for r in recs:
        s = summ.setdefault(r["path"], dict(n=0))
        s["n"] += 1
        for key in ("contract_ratio", "ratio", "rel_rms", "max_abs", "need_atol_rtol2e3", "need_atol_rtol1e4"):
            if key in r:
                if r[key] >= s.get("max_" + key, -1):
                    s["max_" + key] = r[key]
                    s["worst_" + key + "_case"] = r["what"]
        if "psnr" in r:
            if r["psnr"] <= s.get("min_psnr", 1e9):
                s["min_psnr"] = r["psnr"]
        if "peak" in r and "max_abs" in r:
            s["max_abs_rel_peak"] = max(s.get("max_abs_rel_peak", 0.0), r["max_abs"] / r["peak"])
    for path, s in sorted(summ.items()):
        print(path, json.dumps(s))
    tol = json.load(open(TOL))
    tol["measured"] = summ
    json.dump(tol, open(TOL, "w"), indent=2)


if __name__ == "__main__":
    if len(sys.argv) >= 2 and sys.argv[1] == "eref":
        eref()
    elif len(sys.argv) >= 3 and sys.argv[1] == "measured":
        measured(sys.argv[2])
    else:
        print(__doc__)
