"""Shared-memory bank-conflict model used to pick FFT stage orders.

A warp's 16-byte accesses (LDS.128 / STS.128 on the dual-lane 16-byte elements) are served per
quarter-warp: 8 lanes are conflict-free iff they fall in 8 distinct 16-byte bank groups; 8-byte
accesses (float2) are served per half-warp over 16 groups.  For every stage of a plan and every
butterfly leg the script counts the wavefronts per group of lanes; 1.0 = conflict free.
    python scripts/bank_conflict_sim.py
"""


def stages(n, radices):
    out, ns = [], n
    for r in radices:
        out.append((ns, r))
        ns //= r
    return out


def cost(n, radices, lanes):
    tot = cnt = 0
    for ns, r in stages(n, radices):
        L, per = ns // r, n // r
        for id0 in range(0, per, lanes):
            ids = range(id0, min(per, id0 + lanes))
            for q in range(r):
                groups = {}
                for i in ids:
                    b, nn = divmod(i, L)
                    g = (b * ns + nn + q * L) % lanes
                    groups[g] = groups.get(g, 0) + 1
                tot += max(groups.values())
                cnt += 1
    return tot / cnt


if __name__ == "__main__":
    print("column transform, 16-byte dual elements (measured with ncu: 10x10x10x10 -> 1.36, 16x25x25 -> 1.02)")
    for plan in ([10, 10, 10, 10], [16, 25, 25], [20, 20, 25], [25, 16, 25], [25, 25, 16], [10, 10, 20, 5]):
        print(f"  10000 = {'x'.join(map(str, plan)):14s} wavefronts/ideal = {cost(10000, plan, 8):.3f}")
    print("8-byte elements (row kernel, matched-filter blocks, STFT frames)")
    for n, plan in ((10000, [16, 25, 25]), (6000, [16, 25, 15]), (4096, [16, 16, 16]), (2500, [4, 25, 25]), (5000, [8, 25, 25])):
        print(f"  {n:5d} = {'x'.join(map(str, plan)):14s} wavefronts/ideal = {cost(n, plan, 16):.3f}")
