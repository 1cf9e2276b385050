"""Row-sharding plan for the fused dense operator: rank g owns a contiguous, equal-sized block of
rows of A whose boundaries coincide with cone-block boundaries, so A*x slices all-gather in place (rank-major)
and cone blocks never straddle ranks.  Pure host logic (no GPU)."""
from __future__ import annotations


def row_shards(blocks, world):
    """blocks: [(type, len)] back to back.  Returns [(row_offset, n_rows)] per rank, or raises if the product cone
    cannot be cut into `world` equal row ranges on block boundaries."""
    m = sum(ln for _, ln in blocks)
    if world <= 0 or m % world != 0:
        raise ValueError("m=%d does not divide into %d equal shards" % (m, world))
    per = m // world
    bounds = {0}
    acc = 0
    for t, ln in blocks:
        acc += ln
        bounds.add(acc)
        # element-wise cones (Zero=0, RPos=1) may be cut anywhere
        if t in (0, 1):
            bounds.update(range(acc - ln, acc + 1))
    out = []
    for g in range(world):
        lo, hi = g * per, (g + 1) * per
        if lo not in bounds or hi not in bounds:
            raise ValueError("shard boundary %d/%d falls inside a cone block" % (lo, hi))
        out.append((lo, per))
    return out
