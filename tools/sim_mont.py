"""Lane-level Python simulation of csrc/mont.cuh's distributed CIOS (index bookkeeping check)."""
import random
M32 = 0xffffffff

def lookahead(g, p):
    return ((g | p) + g) ^ p

def mont_mul(a, b, n, np0, TPI, L):
    # a,b,n: lists of TPI lists of L limbs
    E = [[0]*L for _ in range(TPI)]; O = [[0]*L for _ in range(TPI)]
    Ec = [0]*TPI; Oc = [0]*TPI
    for u in range(TPI):
        for k in range(L):
            bj = b[u][k]
            sends = [E[t][0] for t in range(TPI)]
            nEs = []; nOs = []; nEcs = []; nOcs = []
            for t in range(TPI):
                recv = sends[t+1] if t < TPI-1 else 0
                top = Ec[t] + recv
                top_lo, top_hi = top & M32, top >> 32
                nE = [0]*L; nO = [0]*L
                s = O[t][0] + E[t][1]; nE[0] = s & M32; cc = s >> 32
                for j in range(0, L-2, 2):
                    v = a[t][j+1]*bj + (E[t][j+2] | (E[t][j+3] << 32)) + cc
                    nO[j] = v & M32; nO[j+1] = (v >> 32) & M32; cc = v >> 64
                v = a[t][L-1]*bj + (top_lo | (top_hi << 32)) + cc
                nO[L-2] = v & M32; nO[L-1] = (v >> 32) & M32; cc = v >> 64
                nOc = cc
                v = a[t][0]*bj + (nE[0] | (O[t][1] << 32))
                nE[0] = v & M32; nE[1] = (v >> 32) & M32; cc = v >> 64
                for j in range(2, L, 2):
                    v = a[t][j]*bj + (O[t][j] | (O[t][j+1] << 32)) + cc
                    nE[j] = v & M32; nE[j+1] = (v >> 32) & M32; cc = v >> 64
                nEc = Oc[t] + cc
                assert nEc <= M32
                nEs.append(nE); nOs.append(nO); nEcs.append(nEc); nOcs.append(nOc)
            q = (nEs[0][0] * np0) & M32
            for t in range(TPI):
                nE, nO = nEs[t], nOs[t]
                cc = 0
                for j in range(0, L, 2):
                    v = n[t][j]*q + (nE[j] | (nE[j+1] << 32)) + cc
                    nE[j] = v & M32; nE[j+1] = (v >> 32) & M32; cc = v >> 64
                nEcs[t] += cc
                cc = 0
                for j in range(0, L, 2):
                    v = n[t][j+1]*q + (nO[j] | (nO[j+1] << 32)) + cc
                    nO[j] = v & M32; nO[j+1] = (v >> 32) & M32; cc = v >> 64
                nOcs[t] += cc
                assert nEcs[t] <= M32 and nOcs[t] <= M32
            assert nEs[0][0] == 0
            E, O, Ec, Oc = nEs, nOs, nEcs, nOcs
    # final
    sends = [E[t][0] for t in range(TPI)]
    r = [[0]*L for _ in range(TPI)]; ov = [0]*TPI
    for t in range(TPI):
        recv = sends[t+1] if t < TPI-1 else 0
        top = Ec[t] + recv
        top_lo, top_hi = top & M32, top >> 32
        cc = 0
        for k in range(L):
            addend = E[t][k+1] if k < L-1 else top_lo
            v = O[t][k] + addend + cc
            r[t][k] = v & M32; cc = v >> 32
        ov[t] = Oc[t] + top_hi + cc
        assert ov[t] <= M32
    return resolve_reduce(r, ov, n, TPI, L)

def resolve_reduce(r, ov, n, TPI, L):
    g = p = 0
    for t in range(TPI):
        inn = ov[t-1] if t > 0 else 0
        val = sum(r[t][k] << (32*k) for k in range(L)) + inn
        cy = val >> (32*L); val &= (1 << (32*L)) - 1
        r[t] = [(val >> (32*k)) & M32 for k in range(L)]
        if cy: g |= 1 << t
        if val == (1 << (32*L)) - 1: p |= 1 << t
        assert cy <= 1
    assert g & p == 0
    ci = lookahead(g, p)
    for t in range(TPI):
        cin = (ci >> t) & 1
        val = (sum(r[t][k] << (32*k) for k in range(L)) + cin) & ((1 << (32*L)) - 1)
        r[t] = [(val >> (32*k)) & M32 for k in range(L)]
    overflow = ov[TPI-1] + ((ci >> TPI) & 1)
    bg = bp = 0; d = []
    for t in range(TPI):
        rv = sum(r[t][k] << (32*k) for k in range(L)); nv = sum(n[t][k] << (32*k) for k in range(L))
        dv = rv - nv
        if dv < 0: bg |= 1 << t; dv += 1 << (32*L)
        if dv == 0: bp |= 1 << t
        d.append(dv)
    assert bg & bp == 0
    bi = lookahead(bg, bp)
    for t in range(TPI):
        d[t] = (d[t] - ((bi >> t) & 1)) & ((1 << (32*L)) - 1)
    final_borrow = (bi >> TPI) & 1
    take = overflow != 0 or final_borrow == 0
    out = 0
    for t in range(TPI):
        v = d[t] if take else sum(r[t][k] << (32*k) for k in range(L))
        out |= v << (32*L*t)
    return out

def split(x, TPI, L):
    return [[(x >> (32*(t*L+k))) & M32 for k in range(L)] for t in range(TPI)]

def test(TPI, L, iters=200):
    S = TPI*L; R = 1 << (32*S)
    rnd = random.Random(1234 + TPI*100 + L)
    for it in range(iters):
        mode = it % 4
        if mode == 0: N = rnd.getrandbits(32*S) | 1 | (1 << (32*S-1))
        elif mode == 1: N = R - 1 - 2*rnd.getrandbits(20)
        elif mode == 2: N = rnd.getrandbits(rnd.randrange(2, 32*S)) | 1 | 2
        else: N = (1 << (32*S-1)) + 1 + 2*rnd.getrandbits(8)
        np0 = (-pow(N, -1, 1 << 32)) & M32
        for a, b in [(rnd.randrange(R), rnd.randrange(N)), (N-1, N-1), (R-1, N-1), (0, 0), (1, 1), (rnd.randrange(N), 0)]:
            got = mont_mul(split(a, TPI, L), split(b, TPI, L), split(N, TPI, L), np0, TPI, L)
            exp = (a*b*pow(R, -1, N)) % N
            assert got == exp, (TPI, L, hex(N), hex(a), hex(b), hex(got), hex(exp))
    print("ok", TPI, L)

if __name__ == "__main__":
    for TPI, L in [(1, 2), (2, 2), (4, 4), (8, 2), (4, 6), (8, 16), (32, 4)]:
        test(TPI, L, 60 if TPI*L > 64 else 200)
