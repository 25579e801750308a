// sp_comm_kernels.cuh -- device side of the multi-GPU entry points (sp_comm.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sp {

// Gathered shards -> database order.  The all-gather buffer holds `world` slots of `slot_rows` rows (one slot per rank, rows in
// that rank's shard order); row_global[slot row] is the database index of the row (-1 = padding).  One CTA per slot row, 16-byte
// copies (ld is a multiple of 64 elements, so rows are 16-byte aligned for u16 and i32).  HBM-bound: reads and writes the
// matrix once (2 x n_total x ld x elem bytes; 102 MB at BASELINE configs[1], ~20 us at the measured 6.4 TB/s).
__global__ void __launch_bounds__(256) comm_rows_to_database_order(const uint4 *__restrict__ src, uint4 *__restrict__ dst,
                                                                   const int32_t *__restrict__ row_global, long long row_vec4) {
    const int g = row_global[blockIdx.x];
    if (g < 0) return;
    const uint4 *s = src + static_cast<long long>(blockIdx.x) * row_vec4;
    uint4 *d = dst + static_cast<long long>(g) * row_vec4;
    for (long long i = threadIdx.x; i < row_vec4; i += blockDim.x) d[i] = s[i];
}

}  // namespace sp
