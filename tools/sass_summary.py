"""SASS evidence of the Blackwell-native instructions (cuobjdump -sass of libgf_attn.so): per kernel the counts of
UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA tensor load / store), UBLKCP (bulk copy),
SYNCS (mbarrier), plus a short excerpt around the first UTCHMMA of three representative kernels.
    python tools/sass_summary.py > profiles/r02/sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "gansformer-reproducibility-challenge_b200", "libgf_attn.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
MNEMS = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "UTCBAR", "UTCATOMSWS")
kernels = collections.OrderedDict()
cur = None
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = []
        continue
    if cur is not None:
        kernels[cur].append(line)
dem = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
print(f"# {lib}: {len(kernels)} kernels; arch line: " + next((l.strip() for l in txt.splitlines() if "arch =" in l), "?"))
print(f"{'kernel':90s} " + " ".join(f"{m:>8s}" for m in MNEMS))
tot = collections.Counter()
for (name, lines), d in zip(kernels.items(), dem):
    cnt = {m: sum(1 for l in lines if re.search(r"\b" + m + r"\b", l)) for m in MNEMS}
    tot.update(cnt)
    if any(cnt.values()):
        print(f"{d[:90]:90s} " + " ".join(f"{cnt[m]:8d}" for m in MNEMS))
print(f"{'TOTAL':90s} " + " ".join(f"{tot[m]:8d}" for m in MNEMS))
for want in ("token_tc_kernel<16, 4, 0, false>", "centroid_tc_kernel<32, 4, 4>", "gemm_tc_kernel"):
    for (name, lines), d in zip(kernels.items(), dem):
        if want in d:
            idx = next((i for i, l in enumerate(lines) if "UTCHMMA" in l), None)
            if idx is None:
                continue
            print(f"\n# excerpt: {d[:110]} (around the first UTCHMMA)")
            for l in lines[max(0, idx - 6): idx + 5]:
                print("   " + l.strip()[:150])
            break
