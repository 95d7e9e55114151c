"""Static shared-memory bank-conflict check of the exchange patterns in
napa_fft3_b200/csrc/napafft_kernels.cuh: for every tile shape, lane mapping (Q = transform index
fastest across lanes, T = position fastest) and layout (interleaved [pos][q] / planar [q][pos])
it replays the 16-byte addresses of every warp instruction and counts 128-byte wavefronts
(a quarter warp is conflict-free iff its 8 addresses fall in 8 distinct 16-byte bank groups).
Run: python tools/bank_check.py      (exits non-zero on any conflict)"""
import sys

PLANS = {4: [4], 5: [3, 2], 6: [3, 3], 7: [4, 3], 8: [4, 4], 9: [4, 4, 1], 10: [4, 4, 2], 11: [4, 4, 3],
         12: [4, 4, 4], 13: [4, 4, 3, 2]}
SMALL_TILE_MAXL = 8


def log2pts(l):
    return 11 if l <= SMALL_TILE_MAXL else (12 if l <= 12 else 13)


def wavefronts(addrs16):
    total = 0
    for qw in range(0, len(addrs16), 8):
        groups = {}
        for a in addrs16[qw:qw + 8]:
            groups.setdefault(a % 8, set()).add(a)
        total += max(len(v) for v in groups.values())
    return total


def ids(tid, l, m):
    lc = log2pts(l) - l
    lt = l - 4
    if m == "Q":
        return tid >> lc, tid & ((1 << lc) - 1)          # t, q
    return tid & ((1 << lt) - 1), tid >> lt


def planar_params(l):
    lc = log2pts(l) - l
    T = 1 << (l - 4)
    padsh = PLANS[l][0]
    if T >= 8:
        skew = 1 if lc >= 3 else (2 if lc == 2 else (4 if lc == 1 else 0))
    else:
        skew = 8 // T if T > 1 else 0
    return padsh, (1 << l) + ((1 << l) >> padsh) + skew


def phys(l, layout, pos, q):
    lc = log2pts(l) - l
    if layout == "I":
        padsh = 31 if lc >= 3 else 4
        return ((pos + (pos >> padsh)) << lc) + q
    padsh, pitch = planar_params(l)
    return q * pitch + pos + (pos >> padsh)


def rphys(l, pos, q):
    """reorder buffer: planar, XOR swizzle of the low 3 bits with the reversed top 3 bits"""
    top = pos >> (l - 3)
    r3 = ((top & 1) << 2) | (top & 2) | (top >> 2)
    return q * (1 << l) + (pos ^ r3)


def rev(x, w):
    return int(format(x, "0%db" % w)[::-1], 2) if w else 0


def check(l, lm, sm):
    L, T = 1 << l, 1 << (l - 4)
    nthreads = 1 << (log2pts(l) - 4)
    layout = "I" if (lm == "Q" and sm == "Q") else "P"
    worst_w = worst_r = 4
    lgns = 0
    radices = PLANS[l]
    for st, lr in enumerate(radices[:-1]):
        R, NB, NS = 1 << lr, 16 >> lr, 1 << lgns
        wm = lm if st == 0 else sm
        for w0 in range(0, nthreads, 32):
            tids = range(w0, min(w0 + 32, nthreads))
            for u in range(NB):
                for s in range(R):
                    addrs = []
                    for tid in tids:
                        t, q = ids(tid, l, wm)
                        jb = t + u * T
                        base = ((jb >> lgns) << (lgns + lr)) + (jb & (NS - 1))
                        addrs.append(phys(l, layout, base + s * NS, q))
                    worst_w = max(worst_w, wavefronts(addrs) * 32 // len(addrs))
            for m in range(16):
                addrs = []
                for tid in tids:
                    t, q = ids(tid, l, sm)
                    addrs.append(phys(l, layout, t + m * T, q))
                worst_r = max(worst_r, wavefronts(addrs) * 32 // len(addrs))
        lgns += lr
    return worst_w, worst_r


def check_reorder(l):
    T = 1 << (l - 4)
    nthreads = 1 << (log2pts(l) - 4)
    worst = 4
    for w0 in range(0, nthreads, 32):
        tids = range(w0, min(w0 + 32, nthreads))
        for m in range(16):
            nat = [rphys(l, ids(tid, l, "T")[0] + m * T, ids(tid, l, "T")[1]) for tid in tids]
            rv = [rphys(l, rev(ids(tid, l, "T")[0] + m * T, l), ids(tid, l, "T")[1]) for tid in tids]
            worst = max(worst, wavefronts(nat), wavefronts(rv))
    return worst


def check_dense(l):
    """dense-rows image [q][L + 1] of the L < 128 single-pass kernels: coalesced side (thread e -> element e)
    and transform side (lane map Q, position t + m*T, natural or bit-reversed)"""
    L, T, C = 1 << l, 1 << (l - 4), 1 << (log2pts(l) - l)
    worst_c = worst_t = 0
    tot_c = n_c = 0
    for w0 in range(0, 128, 32):
        for j in range(16):
            addrs = [((w0 + i + 128 * j) >> l) * (L + 1) + ((w0 + i + 128 * j) & (L - 1)) for i in range(32)]
            wf = wavefronts(addrs)
            worst_c = max(worst_c, wf); tot_c += wf; n_c += 1
        for m in range(16):
            for bitrev in (False, True):
                addrs = []
                for tid in range(w0, w0 + 32):
                    t, q = ids(tid, l, "Q")
                    pos = t + m * T
                    addrs.append(q * (L + 1) + (rev(pos, l) if bitrev else pos))
                worst_t = max(worst_t, wavefronts(addrs))
    return worst_c, tot_c / n_c, worst_t


if __name__ == "__main__":
    bad = 0
    for l in (4, 5, 6):
        wc, avg, wt = check_dense(l)
        flag = "" if wt == 4 and avg < 4.6 else "   <-- CONFLICT"
        bad += flag != ""
        print("L=2^%-2d dense-rows image: coalesced side %.2f wavefronts/instr on average (worst %d, rows straddled), transform side %d (ideal 4)%s" % (l, avg, wc, wt, flag))
    for l in range(5, 14):
        for lm, sm in ((("Q", "Q"),) if l < 7 else (("Q", "Q"), ("T", "T"), ("T", "Q"))):   # L < 128 only ever uses (Q, Q)
            w, r = check(l, lm, sm)
            flag = "" if (w, r) == (4, 4) else "   <-- CONFLICT"
            bad += flag != ""
            print("L=2^%-2d C=%-3d plan=%-16s load map %s store map %s: STS wavefronts/instr %d LDS %d (ideal 4)%s" % (
                l, 1 << (log2pts(l) - l), [1 << x for x in PLANS[l]], lm, sm, w, r, flag))
        if l >= 7:
            w = check_reorder(l)
            flag = "" if w == 4 else "   <-- CONFLICT"
            bad += flag != ""
            print("L=2^%-2d bit-reversal reorder buffer (T map, XOR swizzle): %d wavefronts/instr%s" % (l, w, flag))
    sys.exit(1 if bad else 0)
