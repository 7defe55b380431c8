// Similarity scan of semantic query expansion (reference: SemanticIndex::most_similar_to_vec,
// src/semantic_embedding.cpp:104-146, the loop `sim = dot(qvec, v, dim); if (sim < min_sim) continue;`).
//
// The reference scans every stored vector once per query vector.  Here the L2-normalised vectors live in HBM
// transposed (vT[i][row]), one thread owns one row and carries up to kChunk query vectors at a time, so the
// loads are coalesced and each stored value is read once per chunk.  The dot product is evaluated exactly as
// the reference's `dot` does (src/semantic_embedding.cpp:11-15): s = s + (q[i] * v[i]) for i = 0..dim-1, one
// rounding per multiply and per add, no FMA — sims are bit-identical to the CPU's, which the heap replay on
// the host (host/semantic.hpp select_topk) depends on.  No tensor cores: a split-precision GEMM could not
// reproduce those roundings.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nsb {

constexpr int kSemChunk = 8;  // query vectors per pass over the stored vectors

// out_count[m] counts every survivor of query vector m; the first `cap` are stored (any order).
__global__ void __launch_bounds__(256) cosine_scan_kernel(const float* __restrict__ vT, uint32_t rows, uint32_t dim,
                                                          const float* __restrict__ q, uint32_t M, float min_sim,
                                                          uint32_t cap, uint32_t* __restrict__ out_rows,
                                                          float* __restrict__ out_sims, uint32_t* __restrict__ out_count) {
    extern __shared__ float qs[];  // [kSemChunk][dim]
    const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    for (uint32_t m0 = 0; m0 < M; m0 += kSemChunk) {
        const uint32_t mc = min((uint32_t)kSemChunk, M - m0);
        __syncthreads();
        for (uint32_t t = threadIdx.x; t < mc * dim; t += blockDim.x) qs[t] = q[(size_t)m0 * dim + t];
        __syncthreads();
        if (row < rows) {
            float acc[kSemChunk];
#pragma unroll
            for (int j = 0; j < kSemChunk; j++) acc[j] = 0.0f;
            for (uint32_t i = 0; i < dim; i++) {
                const float v = vT[(size_t)i * rows + row];
#pragma unroll
                for (int j = 0; j < kSemChunk; j++)
                    if ((uint32_t)j < mc) acc[j] = __fadd_rn(acc[j], __fmul_rn(qs[j * dim + i], v));
            }
#pragma unroll
            for (int j = 0; j < kSemChunk; j++) {
                if ((uint32_t)j < mc && !(acc[j] < min_sim)) {  // the reference skips `sim < min_sim`
                    const uint32_t at = atomicAdd(out_count + m0 + j, 1u);
                    if (at < cap) {
                        out_rows[(size_t)(m0 + j) * cap + at] = row;
                        out_sims[(size_t)(m0 + j) * cap + at] = acc[j];
                    }
                }
            }
        }
    }
}

}  // namespace nsb
