"""CPU check of the dual-chain formulation of the fast lane decoder (scan_kernels.cu, BYDB_EXP_DUAL): decoding a lane\nas two 16-byte halves with separate state + one head correction gives exactly the single-chain result."""
import random
M32 = 0xffffffff
def s32(x):
    x &= M32
    return x - (1<<32) if x >> 31 else x
def zz(u): return (u >> 1) ^ -(u & 1)
INT_MAX, INT_MIN = 0x7fffffff, -0x80000000
def single(bs, aw):
    """reference semantics of fast_lane_decode (full chunk): bytes bs[0..32), aw = active-bit window of this lane's values"""
    accv = 0; sh = 0; P = 0; sumP = 0; mn = INT_MAX; mx = INT_MIN
    for b in bs:
        accv |= (b & 0x7f) << sh; sh += 7
        if b < 0x80:
            P += zz(accv)
            if aw & 1:
                sumP += P; mn = min(mn, P); mx = max(mx, P)
            aw >>= 1; accv = 0; sh = 0
    return accv, sh, P, sumP, mn, mx
def head_delta(first_bytes, term, prev_acc, prev_sh):
    fp = (term & -term).bit_length() - 1
    hx = 0
    for k in range(fp + 1): hx |= (first_bytes[k] & 0x7f) << (7 * k)
    full = prev_acc | (hx << prev_sh)
    return zz(full) - zz(hx)
def dual(bs, aw):
    A, B = bs[:16], bs[16:]
    termA = sum(1 << i for i, b in enumerate(A) if b < 0x80)
    termB = sum(1 << i for i, b in enumerate(B) if b < 0x80)
    nA = bin(termA).count("1"); nB = bin(termB).count("1")
    awA = aw & ((1 << nA) - 1); awB = (aw >> nA) & ((1 << nB) - 1)
    accA, shA, PA, sA, mnA, mxA = single(A, awA)
    accB, shB, PB, sB, mnB, mxB = single(B, awB)
    cntB = bin(awB).count("1")
    if nB > 0 and shA != 0:
        d = head_delta(B, termB, accA, shA)
        PB += d; sB += d * cntB
        if cntB: mnB += d; mxB += d
    assert nB > 0 and nA > 0
    P = PA + PB
    sumP = sA + cntB * PA + sB
    mn, mx = mnA, mxA
    if cntB:
        mn = min(mn, PA + mnB); mx = max(mx, PA + mxB)
    return accB, shB, P, sumP, mn, mx
rng = random.Random(5)
def rand_window():
    # narrow varints (1..3 bytes), window may start/end mid-varint; the leading partial acts like lane-start (own-bytes-only)
    out = []
    while len(out) < 40:
        L = rng.choice([1, 1, 2, 2, 2, 3])
        u = rng.randrange(1 << (7 * L))
        for k in range(L):
            out.append(((u >> (7 * k)) & 0x7f) | (0x80 if k < L - 1 else 0))
    s = rng.randrange(0, 4)
    return out[s:s + 32]
bad = 0
for it in range(200000):
    bs = rand_window()
    # skip windows violating the narrow check inside the lane (3 consecutive continuation bytes cannot happen by construction)
    aw = rng.getrandbits(32) if rng.random() < 0.7 else rng.choice([0, M32])
    a = single(bs, aw); b = dual(bs, aw)
    # single-chain: a leading continuation run at the lane start is decoded "own bytes only" in both formulations
    if a != b:
        bad += 1
        if bad < 5: print("MISMATCH", bs, hex(aw), a, b)
print("mismatches", bad)
