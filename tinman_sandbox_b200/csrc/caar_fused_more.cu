// caar_fused_more.cu — fused kernel instances (caar_fused_kernel.cuh) for the level counts beyond the reference's two
// configurations: every multiple of 8 up to 64 as one CTA per element, 80 / 96 / 112 / 120 on CTA clusters.
// Level counts in between run on the next larger instance with padding levels (instance_for()).
#include "caar_fused_kernel.cuh"

namespace caar {

cudaError_t launch_fused_more(const KernelArgs& a, cudaStream_t s) {
  switch (instance_for(a.nlev)) {
    case 8: return launch_nlev<8>(a, s);
    case 16: return launch_nlev<16>(a, s);
    case 24: return launch_nlev<24>(a, s);
    case 32: return launch_nlev<32>(a, s);
    case 40: return launch_nlev<40>(a, s);
    case 48: return launch_nlev<48>(a, s);
    case 56: return launch_nlev<56>(a, s);
    case 64: return launch_nlev<64>(a, s);
    case 80: return launch_nlev<80>(a, s);
    case 96: return launch_nlev<96>(a, s);
    case 112: return launch_nlev<112>(a, s);
    case 120: return launch_nlev<120>(a, s);
  }
  return cudaErrorInvalidValue;
}

}  // namespace caar
