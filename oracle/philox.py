"""CPU restatement of the attention-dropout mask of the kernels (csrc/gf_common.cuh: philox4x32_10 / dropout_mult4).

TEST INFRASTRUCTURE ONLY (see oracle/bipartite.py).  Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as
1, 2, 3", SC'11) with counter (token, column block | step << 8, salt ^ (step >> 24), 0x5eed) and key (seed lo, seed hi); column j
of token t is kept iff word (j % 4) of block j // 4 is >= round(p * 2^32); kept entries carry 1 / (1 - p)."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """All arguments uint64 arrays holding 32-bit values (broadcastable); returns four uint64 arrays of 32-bit words."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & MASK for c in (c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)) & MASK, lo1, (hi0 ^ c3 ^ np.uint64(k1)) & MASK, lo0
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def dropout_mult(p: float, seed: int, step: int, salt: int, tokens: int, KP: int) -> np.ndarray:
    """[tokens, KP] float32 multipliers (0 or 1/(1-p)), identical to gf_attn_dropout_mask."""
    thr = int(p * 4294967296.0 + 0.5)
    thr = min(max(thr, 1), 4294967295)
    scale = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
    tok = np.arange(tokens, dtype=np.uint64)[:, None]
    q = np.arange(KP // 4, dtype=np.uint64)[None, :]
    c1 = (q | np.uint64((step << 8) & 0xFFFFFFFF)) & MASK
    c2 = np.uint64((salt ^ (step >> 24)) & 0xFFFFFFFF)
    w = philox4x32_10(tok + np.zeros_like(q), c1 + np.zeros_like(tok), c2 + np.zeros_like(tok + q), np.uint64(0x5EED) + np.zeros_like(tok + q),
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    words = np.stack(w, axis=2).reshape(tokens, KP)                 # column j = word j % 4 of block j // 4
    return np.where(words >= np.uint64(thr), scale, np.float32(0.0)).astype(np.float32)
