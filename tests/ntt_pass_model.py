"""Pure-Python model of the multi-pass NTT schedule used by csrc/ntt.cu.

Not a test by itself and not product code: it restates, index for index, the
tile / digit / twiddle-exponent arithmetic of `ntt_pass_kernel` so that the
schedule can be checked on the CPU against the oracle (tests/test_ntt_model.py)
before any GPU time is spent.  Keep it in lock-step with ntt.cu.
"""
from __future__ import annotations

from typing import List, Sequence

import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.bn256 import R_MOD

COLS = 8
MAX_S = 9


def plan_digits(k: int) -> List[int]:
    """Digit widths, top digit first; mirrors ntt_plan() in ntt.cu."""
    if k <= MAX_S:
        return [k]
    npass = -(-k // MAX_S)
    base, rem = divmod(k, npass)
    sw = [base + (1 if i < rem else 0) for i in range(npass)]
    return sw


def bitrev(x: int, bits: int) -> int:
    r = 0
    for _ in range(bits):
        r = (r << 1) | (x & 1)
        x >>= 1
    return r


def digit_reverse(hi: int, sw: Sequence[int], nd: int) -> int:
    """hi holds kappa_1 (top) .. kappa_nd (bottom) in position order;
    returns K_low = kappa_1 + kappa_2 << s1 + ..."""
    K = 0
    for q in range(nd - 1, -1, -1):
        d = hi & ((1 << sw[q]) - 1)
        hi >>= sw[q]
        K += d << sum(sw[:q])
    return K


def run_pass(buf_in, buf_out, tw, k, sw, p, n_in=None, pre=None, post=None, n_out=None):
    """One launch of ntt_pass_kernel: every tile, every group, every unit."""
    r = R_MOD
    n = 1 << k
    npass = len(sw)
    s = sw[p]
    W = sum(sw[:p])
    o = k - W - s
    R = 1 << s
    last = p == npass - 1
    first = p == 0
    if n_in is None:
        n_in = n
    if n_out is None:
        n_out = n
    ncols = COLS if npass > 1 else 1
    ntiles = n // (R * ncols)
    for t in range(ntiles):
        if not last:
            lo8 = t & ((1 << (o - 3)) - 1)
            hi = t >> (o - 3)
            in_base = (hi << (k - W)) + (lo8 << 3)
            in_rs, in_cs = 1 << o, 1
            klow0, klow_cs = digit_reverse(hi, sw, p), 0
            out_base, out_rs = in_base, in_rs
        elif npass == 1:
            in_base, in_rs, in_cs = 0, 1, 0
            klow0, klow_cs = 0, 0
            out_base, out_rs = 0, 1
        else:
            s1 = sw[0]
            b1 = t & ((1 << (s1 - 3)) - 1)
            rest_pos = t >> (s1 - 3)
            klow_rest = digit_reverse(rest_pos, sw[1:], p - 1)
            in_base = ((b1 * 8) << (k - s1)) + (rest_pos << s)
            in_cs, in_rs = 1 << (k - s1), 1
            klow0, klow_cs = b1 * 8 + (klow_rest << s1), 1
            out_base, out_rs = klow0, 1 << W
        tile = {}
        # load (first group reads global; modelled as a full tile load)
        for row in range(R):
            for col in range(ncols):
                idx = in_base + row * in_rs + col * in_cs
                x = buf_in[idx] if idx < n_in else 0
                if first and pre is not None:
                    x = x * pre[idx % len(pre)] % r
                tile[(row, col)] = x
        ngroups = (s + 1) // 2
        for g in range(ngroups):
            u = 2 * g + 1
            if u + 1 <= s:
                lbbits = s - u - 1
                D = 1 << lbbits
                for col in range(ncols):
                    klow = klow0 + col * klow_cs
                    for q in range(R // 4):
                        lb = q & (D - 1)
                        hb = q >> lbbits
                        r0 = (hb << (lbbits + 2)) | lb
                        kl = bitrev(hb, u - 1)
                        base = klow + (kl << W)
                        e_u = base << (o + s - u)
                        e1 = base << (o + s - u - 1)
                        e2 = (base + (1 << (W + u - 1))) << (o + s - u - 1)
                        x = [tile[(r0 + m * D, col)] for m in range(4)]
                        cu, c1, c2 = tw[e_u], tw[e1], tw[e2]
                        t2 = x[2] * cu % r
                        t3 = x[3] * cu % r
                        x[0], x[2] = (x[0] + t2) % r, (x[0] - t2) % r
                        x[1], x[3] = (x[1] + t3) % r, (x[1] - t3) % r
                        t1 = x[1] * c1 % r
                        t3 = x[3] * c2 % r
                        x[0], x[1] = (x[0] + t1) % r, (x[0] - t1) % r
                        x[2], x[3] = (x[2] + t3) % r, (x[2] - t3) % r
                        for m in range(4):
                            tile[(r0 + m * D, col)] = x[m]
            else:
                # single radix-2 stage u == s, distance 1
                for col in range(ncols):
                    klow = klow0 + col * klow_cs
                    for hb in range(R // 2):
                        r0 = hb << 1
                        kl = bitrev(hb, u - 1)
                        base = klow + (kl << W)
                        e_u = base << (o + s - u)
                        a, b = tile[(r0, col)], tile[(r0 + 1, col)]
                        tt = b * tw[e_u] % r
                        tile[(r0, col)], tile[(r0 + 1, col)] = (a + tt) % r, (a - tt) % r
        # store
        for row in range(R):
            kap = bitrev(row, s)
            for col in range(ncols):
                x = tile[(row, col)]
                if last:
                    K = out_base + kap * out_rs + col
                    if post is not None:
                        x = x * post[K % len(post)] % r
                    if K < n_out:
                        buf_out[K] = x
                else:
                    buf_out[out_base + kap * out_rs + col] = x


def ntt_model(a: Sequence[int], omega: int, k: int, sw=None, n_in=None, pre=None, post=None,
              n_out=None) -> List[int]:
    n = 1 << k
    if sw is None:
        sw = plan_digits(k)
    assert sum(sw) == k
    tw = [1] * max(n // 2, 1)
    for i in range(1, n // 2):
        tw[i] = tw[i - 1] * omega % R_MOD
    src = list(a) + [0] * (n - len(a))
    scratch = [0] * n
    out = [0] * (n if n_out is None else n_out)
    for p in range(len(sw)):
        last = p == len(sw) - 1
        dst = out if last else scratch
        run_pass(src, dst, tw, k, sw, p, n_in=n_in if p == 0 else None, pre=pre, post=post,
                 n_out=n_out)
        src = scratch
    return out
