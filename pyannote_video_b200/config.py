"""Build-time decisions that were settled by measurement on B200."""

# how taps share an A slab in shared memory: "tap" (one TMA slab per tap, always
# swizzle-atom aligned), "row" (taps of one filter row share a slab, descriptors start at
# arbitrary row offsets) or "all".
SRGEMM_GROUP = "row"   # validated on B200 (gpurun #1): descriptors at arbitrary row offsets read TMA-swizzled slabs correctly
# UMMA descriptor base_offset policy: 0 -> always 0, 1 -> (smem_addr >> 7) & 7
SRGEMM_DESC_MODE = 0
