"""CPU simulation of scan_kernels.cu::delta_page_fast (lane decomposition, head correction, scans)
against a straightforward decode.  Development aid: validates the warp algorithm's arithmetic without a GPU."""
import numpy as np


def zz(u):
    return (u >> 1) ^ -(u & 1)


def enc(vals):
    out = bytearray()
    for v in vals:
        u = (v << 1) ^ (v >> 63)
        u &= (1 << 64) - 1
        while u > 0x7f:
            out.append(0x80 | (u & 0x7f))
            u >>= 7
        out.append(u)
    return bytes(out)


def sim(body, pstart, count, first, active, r1=None):
    """returns (rc, sum, mn, mx, cnt) like the kernel; active: bool per row.
    r1: last row that can be active -> the BYDB_EXP_EARLYSTOP variant stops after the first chunk that passed it."""
    LB = 32  # bytes per lane (kFastLaneBytes)
    total = ((pstart + len(body)) + 15) & ~15
    buf = bytes(pstart) + body + bytes(total - pstart - len(body))
    pend = pstart + len(body)
    S = 0; mn = None; mx = None; cnt = 0
    if active[0]:
        S += first; mn = mx = first; cnt += 1
    V0 = first; carry_acc = 0; carry_sh = 0; row_base = 1
    nchunks = (total + 32 * LB - 1) // (32 * LB)
    for c in range(nchunks):
        lanes = []
        for lane in range(32):
            o = c * 32 * LB + lane * LB
            w = buf[o:o + LB] if o < total else bytes(LB)
            w = w + bytes(LB - len(w))
            lo = min(max(pstart - o, 0), LB); hi = min(max(pend - o, 0), LB)
            valid = ((1 << hi) - 1) & ~((1 << lo) - 1)
            msb = sum(((w[j] >> 7) & 1) << j for j in range(LB))
            term = valid & ~msb; cont = valid & msb
            lanes.append(dict(w=w, lo=lo, hi=hi, valid=valid, term=term, cont=cont))
        # wide check
        wide = False
        for lane, L in enumerate(lanes):
            term, lo, hi = L['term'], L['lo'], L['hi']
            if term:
                first_t = (term & -term).bit_length() - 1
                last_t = term.bit_length() - 1
                lead = first_t - lo; trail = hi - 1 - last_t
            else:
                lead = trail = hi - lo
            L['lead'], L['trail'] = lead, trail
        for lane, L in enumerate(lanes):
            tp = lanes[lane - 1]['trail'] if lane else carry_sh // 7
            c_ = L['cont']
            if (c_ & (c_ >> 1) & (c_ >> 2)) or tp + L['lead'] > 2:
                wide = True
        if wide:
            return (1, None, None, None, None)
        # per lane decode
        n_in = 0
        for L in lanes:
            L['n'] = bin(L['term']).count('1')
            n_in += L['n']; L['n_in'] = n_in
        for lane, L in enumerate(lanes):
            row0 = row_base + L['n_in'] - L['n']
            aw = 0
            for i in range(L['n']):
                if active[row0 + i]:
                    aw |= 1 << i
            acc = 0; sh = 0; kbit = 1; P = 0; sumP = 0; minP = 2**31 - 1; maxP = -2**31; head_v = 0; head_x = 0
            for j in range(LB):
                b = L['w'][j]
                if (L['valid'] >> j) & 1:
                    acc |= (b & 0x7f) << sh; sh += 7
                if (L['term'] >> j) & 1:
                    v = zz(acc)
                    if kbit == 1:
                        head_x, head_v = acc, v
                    P += v
                    if aw & kbit:
                        sumP += P; minP = min(minP, P); maxP = max(maxP, P)
                    kbit <<= 1; acc = 0; sh = 0
            L.update(acc=acc, sh=sh, P=P, sumP=sumP, minP=minP, maxP=maxP, head_v=head_v, head_x=head_x, aw=aw)
        new_carry = (lanes[31]['acc'], lanes[31]['sh'])
        for lane, L in enumerate(lanes):
            pa, ps = (lanes[lane - 1]['acc'], lanes[lane - 1]['sh']) if lane else (carry_acc, carry_sh)
            cntA = bin(L['aw']).count('1')
            if L['n'] > 0 and ps != 0:
                term = L['term']; fp = (term & -term).bit_length() - 1
                assert fp <= 2 and L['lo'] == 0
                xb = int.from_bytes(L['w'][:4], 'little') & (0xffffff >> (8 * (2 - fp)))
                hx = (xb & 0x7f) | ((xb >> 1) & 0x3f80) | ((xb >> 2) & 0x1fc000)
                assert hx == L['head_x']
                dlt = zz(pa | (hx << ps)) - zz(hx)
                L['P'] += dlt; L['sumP'] += dlt * cntA
                if cntA:
                    L['minP'] += dlt; L['maxP'] += dlt
            L['cntA'] = cntA
        carry_acc, carry_sh = new_carry
        s_in = 0
        for L in lanes:
            s_in += L['P']; L['s_in'] = s_in
        for L in lanes:
            base = V0 + L['s_in'] - L['P']
            if L['cntA']:
                S += base * L['cntA'] + L['sumP']
                a, b = base + L['minP'], base + L['maxP']
                mn = a if mn is None else min(mn, a); mx = b if mx is None else max(mx, b)
                cnt += L['cntA']
        V0 += lanes[31]['s_in']; row_base += lanes[31]['n_in']
        if r1 is not None and row_base > r1 and c + 1 < nchunks:
            return (0, S, mn, mx, cnt)            # early stop: nothing after r1 is active, the tail is not decoded
    ok = row_base == count and carry_sh == 0
    return (0 if ok else 2, S, mn, mx, cnt)


def main():
    rng = np.random.default_rng(1)
    for trial in range(300):
        n = int(rng.integers(2, 3000))
        scale = int(rng.choice([30, 60, 5000, 200000, 1 << 19, 1 << 22]))
        deltas = rng.integers(-scale, scale + 1, n - 1).tolist()
        first = int(rng.integers(-10**12, 10**12))
        vals = np.cumsum([first] + deltas).tolist()
        body = enc(deltas)
        pstart = int(rng.integers(0, 16))
        active = (rng.random(n) < rng.choice([1.0, 0.5, 0.05])).tolist()
        rc, S, mn, mx, cnt = sim(body, pstart, n, first, active)
        maxlen = max(len(enc([d])) for d in deltas)
        if maxlen > 3:
            assert rc == 1, (trial, maxlen, rc)
            continue
        assert rc == 0, (trial, rc, maxlen)
        av = [v for v, a in zip(vals, active) if a]
        assert cnt == len(av) and S == sum(av), (trial, cnt, len(av))
        if av:
            assert mn == min(av) and mx == max(av), trial
        # the early-stop experiment: rows after a random r1 are inactive; stopping at the first chunk past r1 changes nothing
        r1 = int(rng.integers(0, n))
        act2 = [a and i <= r1 for i, a in enumerate(active)]
        full = sim(body, pstart, n, first, act2)
        early = sim(body, pstart, n, first, act2, r1=r1)
        assert early[0] == 0 and early[1:] == full[1:], (trial, r1)
    print("fast path simulation ok")


if __name__ == "__main__":
    main()
