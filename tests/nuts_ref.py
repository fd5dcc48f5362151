"""Host restatement of the batched NUTS transition (mcmc-date_b200/csrc/hmc_kernels.cuh) for the tests: the same
Algorithm-3 tree building chain by chain, with the ORACLE's value / gradient and bit-identical Philox4x32-10 uniforms.
Test infrastructure only."""
import math

import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF
DELTA_MAX = 1000.0


def philox4x32_10(counter, key):
    c = list(counter)
    k0, k1 = key
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c[3] ^ k1) & MASK, p0 & MASK]
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c


def uniform(seed, chain, iteration, draw, stream=0):
    c = philox4x32_10((chain, iteration, draw, stream), (seed & MASK, (seed >> 32) & MASK))
    k = ((c[0] >> 5) << 26) | (c[1] >> 6)
    return (k + 0.5) * 2.0 ** -53


def nuts_chain(orc, x_base, theta0, mom0, inv_mass, eps, max_depth, seed, iteration, chain):
    """one NUTS transition of one chain -> (theta, out[7], accept_stat, (depth, n_leap, diverged, n), status)"""

    def evaluate(th):
        out, g, st = orc.eval_grad(orc.from_vector(x_base, th)[None, :])
        return out[0], orc.to_vector(g[0]), int(st[0])

    out0, g0, st0 = evaluate(theta0)
    status = st0
    h0neg = out0[6] - 0.5 * float(np.sum(mom0 * mom0 * inv_mass))
    u = uniform(seed, chain, iteration, 0)
    ud = uniform(seed, chain, iteration, 1)
    draw = 2
    logu = h0neg + math.log(u)
    th_m, out_m = theta0.copy(), out0.copy()
    if not math.isfinite(h0neg):
        return th_m, out_m, 0.0, (0, 0, 0, 1), status
    ends = [dict(th=theta0.copy(), r=mom0.copy(), g=g0.copy()), dict(th=theta0.copy(), r=mom0.copy(), g=g0.copy())]
    direction = 0 if ud < 0.5 else 1
    depth, n, alpha, n_alpha, n_leap, diverged_any = 0, 1, 0.0, 0, 0, 0
    while True:
        # build 2^depth leaves at end `direction`
        v = 1.0 if direction else -1.0
        e = v * eps
        E = ends[direction]
        ck = {}
        n_sub, s_sub, th_c, out_c = 0, True, None, None
        for leaf in range(1 << depth):
            p = E["r"] + 0.5 * e * E["g"]
            th = E["th"] + e * inv_mass * p
            out, g, st = evaluate(th)
            p = p + 0.5 * e * g
            E["th"], E["r"], E["g"] = th, p, g
            status |= st
            hneg = out[6] - 0.5 * float(np.sum(p * p * inv_mass))
            valid = logu <= hneg
            diverged = not (hneg > logu - DELTA_MAX)
            a = math.exp(hneg - h0neg) if not math.isnan(hneg - h0neg) and hneg - h0neg < 700 else (float("inf") if hneg - h0neg >= 700 else float("nan"))
            alpha += min(1.0, a) if a == a else 0.0
            n_alpha += 1
            n_leap += 1
            if valid:
                n_sub += 1
                uu = uniform(seed, chain, iteration, draw)
                draw += 1
                if uu * n_sub < 1.0:
                    th_c, out_c = th.copy(), out.copy()
            turned = False
            if diverged:
                diverged_any = 1
            else:
                idx_max = bin(leaf >> 1).count("1")
                if leaf % 2 == 0:
                    ck[idx_max] = (th.copy(), p.copy())
                else:
                    nsub = 0
                    while (leaf >> nsub) & 1:
                        nsub += 1
                    for k in range(idx_max, idx_max - nsub, -1):
                        cth, cr = ck[k]
                        dth = (th - cth) * inv_mass
                        if v * float(np.sum(dth * cr)) < 0.0 or v * float(np.sum(dth * p)) < 0.0:
                            turned = True
                            break
            if diverged or turned:
                s_sub = False
                break
        if not s_sub:
            break
        uu = uniform(seed, chain, iteration, draw)
        draw += 1
        if n_sub > 0 and uu * n < n_sub:
            th_m, out_m = th_c, out_c
        n += n_sub
        dth = (ends[1]["th"] - ends[0]["th"]) * inv_mass
        turned_main = float(np.sum(dth * ends[0]["r"])) < 0.0 or float(np.sum(dth * ends[1]["r"])) < 0.0
        depth += 1
        if turned_main or depth >= max_depth:
            break
        ud = uniform(seed, chain, iteration, draw)
        draw += 1
        direction = 0 if ud < 0.5 else 1
    return th_m, out_m, (alpha / n_alpha if n_alpha else 0.0), (depth, n_leap, diverged_any, n), status
