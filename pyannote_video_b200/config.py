"""Build-time decisions that were settled by measurement on B200."""
import os

# how taps share an A slab in shared memory: "tap" (one TMA slab per tap, always
# swizzle-atom aligned), "row" (taps of one filter row share a slab, descriptors start at
# arbitrary row offsets) or "all".
SRGEMM_GROUP = "row"   # validated on B200 (gpurun #1): descriptors at arbitrary row offsets read TMA-swizzled slabs correctly
# (gpurun #1 also settled the UMMA descriptor question: base_offset stays 0; the tensor core applies the
# swizzle XOR to absolute shared-memory address bits, exactly like TMA does when it writes the slab.)

# first detector conv input:
#   "fused"    conv1_fused.cu reads the RGBA u8 plane by TMA, normalises every pixel once into a
#              shared-memory pixel-row buffer and lets tcgen05.mma read OVERLAPPING A rows from it
#              (no im2col in HBM or in shared memory).  8 frames 1080p: 1.36 ms
#   "gathered" pack kernel writes kw*3-wide bf16 rows (16 B/pixel), generic srgemm reads them back:
#              1.15 + 1.73 ms
#   "pixrows"  pack kernel writes a bf16 RGBX plane (8 B/pixel), srgemm reads 8-pixel runs through an
#              overlapping-row tensor map (row stride 16 B): slower than "gathered" (L2-bound re-reads)
#   "c12"      csrc/c12.cu: conv1 AND conv2 in one row-streaming strip kernel — the conv1 activations go from TMEM to
#              shared memory in conv2's operand layout and never reach HBM (464 MB per 1080p frame less traffic)
#              [default with DET_CONVS = "rsconv"; any other conv back-end falls back to "fused"]
DET_CONV1 = os.environ.get("PV_DET_CONV1", "c12")    # measured (8 frames 1080p): c12 1.05 ms vs conv1_fused 0.875 + conv2 0.473 ms

# detector conv layers 2..7:
#   "rsconv"   csrc/rsconv.cu: row streaming — every input row is loaded once and multiplied against all filter rows
#              that use it side by side (N up to 5 x 48 = 240 per MMA, accumulators = a ring of TMEM row slots):
#              tensor-bound instead of shared-memory-bound                                                  [default]
#   "detconv"  csrc/detconv.cu: 2-D tiles (8 x 16 outputs), one 4-D TMA patch per tile, every tap's A operand is a
#              UMMA descriptor into the patch (1.9 input pixels read per output), compile-time MMA sequence
#   "srgemm"   the generic 1-D shifted-row GEMM (5.3 input rows read per output, table-driven issue loop)
DET_CONVS = os.environ.get("PV_DET_CONVS", "rsconv")

# embedder convs of levels 4 / 3 (32 and 64 channels, 14 of the 29 convs, 54 % of the FLOPs):
#   "rs"      csrc/rsconv.cu, faces packed side by side in an image row (N = 3 x Cout per MMA)          [default]
#   "srgemm"  the shifted-row GEMM for every conv (round-1 path; N = Cout <= 64 is shared-memory-bound)
EMBED_IMPL = os.environ.get("PV_EMBED_IMPL", "rs")
