// tools/check_tcgen05_layout.cu -- NOT part of the product build.  Host-only cross-check of the hand-written descriptor
// and shared-memory index math of tools/tcgen05_gemm_probe.cu (the round-2 nnet3 GEMM) against CuTe, from the CUTLASS
// headers vendored in this image (no GPU needed: everything evaluated here is host constexpr / layout algebra):
//   * the 32-bit instruction descriptor for kind::tf32, F32 accumulate, M = N = 128, both operands K-major;
//   * the canonical K-major no-swizzle ("INTERLEAVE") layout of a 128 x 32 tf32 operand tile: byte offset of every (row, k);
//   * the shared-memory matrix descriptor fields (leading / stride byte offsets in 16-byte units, version, layout type).
//
//   P=$(python -c "import flashinfer,os;print(os.path.join(os.path.dirname(flashinfer.__file__),'data/cutlass/include'))")
//   nvcc -std=c++17 -I$P -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -o /tmp/chk tools/check_tcgen05_layout.cu && /tmp/chk
//
// Result in the build container (2026-09-23): idesc 0x08200910 both; 0 layout mismatches over 128 x 32; CuTe descriptor
// LBO = 128 (x16 B = 2048 B = one 16-byte K chunk of all 128 rows), SBO = 8 (x16 B = 128 B = 8 rows), version 1, layout 0 --
// the values make_desc(saddr, PANEL_BYTES, 128) of the probe encodes.  (CuTe prints "cast_smem_ptr_to_uint not supported"
// on the host: the start-address field is the only one that needs a device.)
#include <cstdio>
#include <cute/tensor.hpp>
#include <cute/arch/mma_sm100_desc.hpp>
#include <cute/atom/mma_traits_sm100.hpp>
using namespace cute;

int main() {
  const auto d = UMMA::make_instr_desc<tfloat32_t, tfloat32_t, float, 128, 128, UMMA::Major::K, UMMA::Major::K>();
  const uint32_t mine = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // probe: make_idesc()
  printf("instruction descriptor: cute %08x probe %08x %s\n", (uint32_t)d, mine, (uint32_t)d == mine ? "OK" : "MISMATCH");
  const auto lay = tile_to_shape(UMMA::Layout_K_INTER_Atom<tfloat32_t>{}, Shape<_128, _32>{});
  int bad = 0;
  for (int r = 0; r < 128; r++)
    for (int k = 0; k < 32; k++) {
      const int off_cute = (int)lay(r, k) * 4;
      const int off_probe = (k >> 2) * 2048 + r * 16 + (k & 3) * 4;          // probe: tile_off(r, k)
      bad += off_cute != off_probe;
    }
  printf("operand tile layout: %d mismatches over 128 x 32\n", bad);
  alignas(1024) static float buf[128 * 32];
  const auto t = make_tensor(make_smem_ptr(reinterpret_cast<tfloat32_t *>(buf)), lay);
  const auto desc = UMMA::make_umma_desc<UMMA::Major::K>(t);
  printf("matrix descriptor: LBO %u SBO %u (16-byte units) version %u layout %u; probe encodes LBO %u SBO %u version 1 layout 0\n",
         (unsigned)desc.leading_byte_offset_, (unsigned)desc.stride_byte_offset_, (unsigned)desc.version_, (unsigned)desc.layout_type_,
         2048u >> 4, 128u >> 4);
  const bool ok = (uint32_t)d == mine && bad == 0 && desc.leading_byte_offset_ == (2048u >> 4) && desc.stride_byte_offset_ == (128u >> 4) &&
                  desc.version_ == 1 && desc.layout_type_ == 0;
  printf("%s\n", ok ? "ALL OK" : "DIFFERENCES FOUND");
  return ok ? 0 : 1;
}
