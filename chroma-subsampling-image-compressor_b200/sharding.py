"""Multi-GPU sharding of the pixel path -- host-side arithmetic only, no collective.

The reference builds a fresh DUT per image (ImageCompressorTopApp.scala:53), so frames share no
state: rank g of G takes a contiguous slice of the batch (`frame_shard`).  A single frame can also
be cut into row bands (`band_plan`): the only cross-row state is the held chroma of 4:2:0 / 4:1:0
lines (ChromaSubsampler.scala:34-35,57-65), and a band that starts on an aligned row never needs it,
so aligned bands need zero halo.  Output stays sharded; nothing is exchanged on the hot path.
"""
import math


def frame_shard(n_frames, rank, world):
    """[lo, hi) of the frames rank `rank` of `world` processes; sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return n_frames * rank // world, n_frames * (rank + 1) // world


def band_alignment(factor, a, b, chroma_first):
    """Output rows per alignment unit such that a band starting on a unit boundary has no
    dependency on rows above it.  chroma before spatial: input rows must be a multiple of
    lcm(f, vf) -> in output rows lcm(f,vf)/f.  spatial before chroma (f>1): one counter line spans
    f output rows and odd lines replay the line above -> f*vf output rows."""
    vf = 2 if b == 0 else 1
    if chroma_first or factor == 1:
        return math.lcm(factor, vf) // factor
    return factor * vf


def band_plan(out_h, world, factor, a, b, chroma_first):
    """Split `out_h` output rows into `world` aligned bands: list of (out_row0, out_rows).
    e.g. 8K, f=4, 4:2:0, chroma first: unit 1 output row... spatial first: unit 8 rows."""
    unit = band_alignment(factor, a, b, chroma_first)
    units = -(-out_h // unit)
    bands = []
    for r in range(world):
        u0, u1 = units * r // world, units * (r + 1) // world
        r0, r1 = min(u0 * unit, out_h), min(u1 * unit, out_h)
        bands.append((r0, r1 - r0))
    return bands
