// hq_umma.cu -- the tcgen05 lone-gate kernels (hq_umma.cuh) as a translation unit of the library.
#include "hq_umma.cuh"

#include "hq_kernels.h"

namespace hq {

int launch_umma(void* state, unsigned n_qubits, const unsigned* pos, unsigned k, const void* d_operands, void* stream) {
  UmmaPos p;
  for (unsigned i = 0; i < 8; ++i) p.tpos[i] = i < k ? (unsigned char)pos[i] : 0;
  for (unsigned i = 1; i < k; ++i)
    if (pos[i] <= pos[i - 1]) return int(cudaErrorInvalidValue);
  const size_t floats = (size_t(2) << k) * (size_t(2) << k);
  const float4* bhi = static_cast<const float4*>(d_operands);
  const float4* blo = reinterpret_cast<const float4*>(static_cast<const float*>(d_operands) + floats);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  switch (k) {
    case 4: return launch_umma_gate<4>(static_cast<float2*>(state), n_qubits, p, bhi, blo, s);
    case 5: return launch_umma_gate<5>(static_cast<float2*>(state), n_qubits, p, bhi, blo, s);
    case 6: return launch_umma_gate<6>(static_cast<float2*>(state), n_qubits, p, bhi, blo, s);
    default: return int(cudaErrorInvalidValue);
  }
}

}  // namespace hq
