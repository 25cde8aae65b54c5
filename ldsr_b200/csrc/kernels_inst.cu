// kernels_inst.cu -- compiled once per supported padded width: nvcc -DLDSR_PQ=<n>.
#include "kernel_table.h"

#ifndef LDSR_PQ
#error "compile with -DLDSR_PQ=<n>"
#endif

namespace ldsr {
namespace {

constexpr int PQ = LDSR_PQ;

cudaError_t em_prepare(size_t smem_bytes) {
    cudaError_t e = cudaFuncSetAttribute(em_chunk_kernel<PQ, EM_SEG, EM_WARPS, 1>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(em_chunk_kernel<PQ, EM_SEG, EM_WARPS, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)smem_bytes);
}
cudaError_t em_chunk(const EmParams &p, int n_tasks, size_t smem_bytes, cudaStream_t st) {
    if (p.mode == 2)
        em_chunk_kernel<PQ, EM_SEG, EM_WARPS, 2><<<n_tasks, EM_WARPS * 32, smem_bytes, st>>>(p);
    else if (p.mode == 1)
        em_chunk_kernel<PQ, EM_SEG, EM_WARPS, 1><<<n_tasks, EM_WARPS * 32, smem_bytes, st>>>(p);
    else
        em_chunk_kernel<PQ, EM_SEG, EM_WARPS, 0><<<n_tasks, EM_WARPS * 32, 0, st>>>(p);
    return cudaGetLastError();
}
constexpr int MINB = split_minb_for(PQ);
constexpr int MINB_WIDE = 2;
cudaError_t em_split_prepare(size_t smem_bytes) {
    cudaError_t e = cudaFuncSetAttribute(em_split_kernel<PQ, SPLIT_NW, MINB, SPLIT_MSEG, SPLIT_UW>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess || MINB == MINB_WIDE) return e;
    return cudaFuncSetAttribute(em_split_kernel<PQ, SPLIT_NW, MINB_WIDE, SPLIT_MSEG, SPLIT_UW>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
}
cudaError_t em_split_wide(const SplitParams &p, int n_tasks, size_t smem_bytes, cudaStream_t st) {
    em_split_kernel<PQ, SPLIT_NW, MINB_WIDE, SPLIT_MSEG, SPLIT_UW><<<n_tasks, SPLIT_NW * 32, smem_bytes, st>>>(p);
    return cudaGetLastError();
}
cudaError_t em_split(const SplitParams &p, int n_tasks, size_t smem_bytes, cudaStream_t st) {
    em_split_kernel<PQ, SPLIT_NW, MINB, SPLIT_MSEG, SPLIT_UW><<<n_tasks, SPLIT_NW * 32, smem_bytes, st>>>(p);
    return cudaGetLastError();
}
cudaError_t smoother(const SmootherParams &p, cudaStream_t st) {
    smoother_kernel<PQ><<<(p.n_jobs + 63) / 64, 64, 0, st>>>(p);
    return cudaGetLastError();
}
cudaError_t mstep(const MstepParams &p, cudaStream_t st) {
    mstep_kernel<PQ><<<(p.n_fits + 63) / 64, 64, 0, st>>>(p);
    return cudaGetLastError();
}
cudaError_t propagate(const SmootherParams &p, cudaStream_t st) {
    propagate_kernel<PQ><<<(p.n_jobs + 63) / 64, 64, 0, st>>>(p);
    return cudaGetLastError();
}
cudaError_t rep(const RepParams &p, cudaStream_t st) {
    rep_kernel<PQ><<<(p.n_reps + 127) / 128, 128, 0, st>>>(p);
    return cudaGetLastError();
}

const KernelTable table = {PQ, em_prepare, em_chunk, SPLIT_NW, MINB, SPLIT_MSEG, SPLIT_UW, em_split_prepare, em_split, em_split_wide, smoother, mstep, propagate, rep};

} // namespace

#define LDSR_CAT2(a, b) a##b
#define LDSR_CAT(a, b) LDSR_CAT2(a, b)
const KernelTable *LDSR_CAT(kernel_table_pq, LDSR_PQ)() { return &table; }

} // namespace ldsr
