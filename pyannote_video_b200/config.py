"""Build-time decisions that were settled by measurement on B200."""
import os

# how taps share an A slab in shared memory: "tap" (one TMA slab per tap, always
# swizzle-atom aligned), "row" (taps of one filter row share a slab, descriptors start at
# arbitrary row offsets) or "all".
SRGEMM_GROUP = "row"   # validated on B200 (gpurun #1): descriptors at arbitrary row offsets read TMA-swizzled slabs correctly
# (gpurun #1 also settled the UMMA descriptor question: base_offset stays 0; the tensor core applies the
# swizzle XOR to absolute shared-memory address bits, exactly like TMA does when it writes the slab.)

# first detector conv input: "gathered" (pack kernel writes kw*3-wide rows, 16 B/pixel) or
# "pixrows" (conv reads 8-pixel runs of a bf16 RGBX plane in place through an overlapping-row
# tensor map, 8 B/pixel; needs cuTensorMapEncodeTiled to accept a 16-byte row stride)
DET_CONV1 = os.environ.get("PV_DET_CONV1", "gathered")
