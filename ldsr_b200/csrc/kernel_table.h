// kernel_table.h -- host-visible launchers of the per-PQ kernel instantiations.
// Each supported padded input width PQ is compiled as its own translation unit
// (kernels_inst.cu with -DLDSR_PQ=n) and exposes one table of launch functions.
#pragma once
#include "aux_kernels.cuh"
#include "em_kernel.cuh"

namespace ldsr {

#ifndef LDSR_SEG
#define LDSR_SEG 8
#endif
constexpr int EM_SEG = LDSR_SEG; // steps per checkpoint segment (divides 32)
constexpr int EM_WARPS = 4; // warps per CTA (one per SM sub-partition)

struct KernelTable {
    int pq;
    cudaError_t (*em_prepare)(size_t smem_bytes); // opt in to > 48 KB dynamic shared memory
    cudaError_t (*em_chunk)(const EmParams &, int n_tasks, size_t smem_bytes, cudaStream_t);
    cudaError_t (*smoother)(const SmootherParams &, cudaStream_t);
    cudaError_t (*mstep)(const MstepParams &, cudaStream_t);
    cudaError_t (*propagate)(const SmootherParams &, cudaStream_t);
    cudaError_t (*rep)(const RepParams &, cudaStream_t);
};

const KernelTable *kernel_table_for(int pq_needed); // smallest instantiation with pq >= pq_needed

} // namespace ldsr
