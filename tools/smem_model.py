#!/usr/bin/env python
"""Strict wavefront model of the plan-400 kernel's shared-memory accesses, used to choose the lane mapping and layouts.

Rule (measured with tools/micro/smem_rule.cu on B200): a 64-bit access is split into fixed half-warps and a 128-bit access
into fixed quarter-warps; a group costs as many wavefronts as its most-loaded bank pair / bank quad.  The search below
scores, for every candidate (lane -> (worker t, FFT g) mapping for step 1 and step 3, worker permutation, Z row stride and
row layout, power-slab layout, PCM chunk pads), the wavefronts in excess of the ideal for
  the 28 PCM LDS.64 of step 1, the 20 Z STS.128, the 20 Z LDS.128 and the 20 power STS.64 of a pass.
Result (what melspec400_kernel uses): lane = 10 g + t for both steps, chunk pads of 20 words, Z rows [g][t] with a stride
of 33 units, planar power slabs with origins (0, 6, 0) mod 16: everything conflict-free except the power store (3 instead
of 2 wavefronts: with 10 + 6 lanes of two FFTs in a half-warp and worker 0's bin sitting at +10 / -0 no origin works for
both the ascending and the mirrored half).  The previous mapping lane = 3 t + g cost 28 extra wavefronts on the PCM loads."""
import itertools


def lanes_map(kind):
    m = []
    for l in range(32):
        ll = min(l, 29)                      # lanes 30, 31 shadow lane 29
        m.append((ll // 3, ll % 3) if kind == "t3g" else (ll % 10, ll // 10))
    return m


def wf(units, group, mod):
    """wavefronts of one access: `units` = per-lane unit index (8- or 16-byte units), fixed groups of `group` lanes"""
    tot = 0
    for h in range(0, 32, group):
        cnt = {}
        for u in set(units[h:h + group]):    # identical addresses broadcast
            cnt[u % mod] = cnt.get(u % mod, 0) + 1
        tot += max(cnt.values())
    return tot


def pcm_cost(kind1, pads):
    lm = lanes_map(kind1)
    start = [0]
    for k in range(3):
        start.append(start[-1] + 320 + pads[k])
    tot = 0
    for m in range(28):
        units = []
        for (t, g) in lm:
            n = 320 * g + 20 * m + 2 * t      # tile sample of (row m, columns 2t, 2t+1) of FFT g
            units.append((start[n // 320] + n % 320) // 2)
        tot += wf(units, 16, 16)
    return tot                                # ideal 56


def zst_cost(kind1, zi, zg):
    return wf([zi * t + zg * g for (t, g) in lanes_map(kind1)], 8, 8)          # ideal 4


def zld_cost(kind3, pi, zrow, zg):
    return wf([zrow * pi[t] + zg * g for (t, g) in lanes_map(kind3)], 8, 8)    # ideal 4


def pst_cost(kind3, pi, layout, P):
    tot = 0
    for half in (0, 1):                       # slots j < 10 (ascending bins) and j >= 10 (mirrored bins)
        units = []
        for (t, g) in lanes_map(kind3):
            w = pi[t]
            off = (w if w else 10) if half == 0 else -w
            b = 64 + off
            units.append(3 * b + g if layout == "L1" else g * 1024 + P[g] + b)
        tot += wf(units, 16, 16)
    return tot                                # ideal 4


def main():
    perms = [[(a * t + b) % 10 for t in range(10)] for a in (1, 3, 7, 9) for b in range(10)]
    res = []
    for kind1 in ("t3g", "g10t"):
        pc = min((pcm_cost(kind1, p), p) for p in itertools.product(range(0, 33, 4), repeat=3))
        for (zi, zg) in ((3, 1), (1, 10)):
            zs = zst_cost(kind1, zi, zg)
            for kind3 in ("t3g", "g10t"):
                for pi in perms:
                    for zrow in (31, 33, 35, 37):
                        zl = zld_cost(kind3, pi, zrow, zg)
                        layouts = [("L1", None)] + [("L2", (0, a, b)) for a in range(16) for b in range(16)]
                        for layout, P in layouts:
                            ps = pst_cost(kind3, pi, layout, P)
                            excess = (pc[0] - 56) + 20 * (zs - 4) + 20 * (zl - 4) + 10 * (ps - 4)
                            res.append((excess, kind1, pc, (zi, zg), zs, kind3, pi, zrow, zl, layout, P, ps))
    res.sort(key=lambda r: r[0])
    print("excess wavefronts per pass, step-1 mapping, (PCM wavefronts, pads), Z unit strides (pair, g), Z STS, step-3 mapping, "
          "worker permutation, Z row stride, Z LDS, power layout, plane origins, power STS (2 instr.)")
    seen = set()
    for r in res:
        key = (r[1], r[5])
        if key not in seen:
            seen.add(key)
            print(r)


if __name__ == "__main__":
    main()
