"""Static shared-memory bank-conflict check of the exchange patterns in
napa_fft3_b200/csrc/napafft_kernels.cuh (Stages<>::run): for every tile shape it replays the
addresses each warp instruction touches and counts 128-byte wavefronts (a quarter warp of 16-byte
accesses is conflict-free iff its 8 addresses fall in 8 distinct 16-byte bank groups).
Run: python tools/bank_check.py"""
PLANS = {4: [4], 5: [3, 2], 6: [3, 3], 7: [4, 3], 8: [4, 4], 9: [3, 3, 3], 10: [4, 4, 2], 11: [4, 4, 3],
         12: [4, 4, 4], 13: [4, 4, 3, 2]}


def wavefronts(addrs16):
    """addrs16: 32 lane addresses in 16-byte units -> number of 128 B wavefronts for an LDS/STS.128"""
    total = 0
    for qw in range(4):
        lanes = addrs16[8 * qw: 8 * qw + 8]
        groups = {}
        for a in lanes:
            groups.setdefault(a % 8, set()).add(a)
        total += max(len(v) for v in groups.values())
    return total


def check(l, padshift=None):
    L = 1 << l
    T = L // 16
    lc = 12 - l if l <= 12 else 0
    C = 1 << lc
    if padshift is None:
        padshift = 31 if lc >= 3 else 4
    nthreads = T * C
    phys = lambda pos, q: ((pos + (pos >> padshift)) << lc) + q
    worst_r = worst_w = 4
    lgns = 0
    radices = PLANS[l]
    for st, lr in enumerate(radices[:-1]):
        R, NB, NS = 1 << lr, 16 >> lr, 1 << lgns
        for w0 in range(0, nthreads, 32):
            tids = range(w0, w0 + 32)
            for u in range(NB):
                for s in range(R):
                    addrs = []
                    for tid in tids:
                        q, t = tid & (C - 1), tid >> lc
                        jb = t + u * T
                        base = ((jb >> lgns) << (lgns + lr)) + (jb & (NS - 1))
                        addrs.append(phys(base + s * NS, q))
                    worst_w = max(worst_w, wavefronts(addrs))
            for m in range(16):
                addrs = [phys((tid >> lc) + m * T, tid & (C - 1)) for tid in tids]
                worst_r = max(worst_r, wavefronts(addrs))
        lgns += lr
    return worst_w, worst_r


if __name__ == "__main__":
    for l in range(5, 14):
        w, r = check(l)
        print("L=2^%-2d C=%-3d plan=%s  worst STS wavefronts/instr %d (ideal 4), LDS %d" % (
            l, 1 << (12 - l if l <= 12 else 0), [1 << x for x in PLANS[l]], w, r))
