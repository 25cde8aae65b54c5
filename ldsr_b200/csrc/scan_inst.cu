// scan_inst.cu -- instantiations and launch sequence of the general-state-dimension smoother.
#include <cstdlib>

#include "scan_kernels.cuh"

namespace ldsr {
namespace {

template <class K> cudaError_t opt_in(K kernel, size_t smem) { // above the 48 KB default
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}
#define LDSR_TRY(x)                      \
    do {                                 \
        cudaError_t e_ = (x);            \
        if (e_ != cudaSuccess) return e_; \
    } while (0)

template <int D> cudaError_t run(const ScanParams &P, cudaStream_t st) {
    const long long nt = (long long)P.n_fits * P.T;
    // one thread per (fit, chunk), long dependent chains: with few chunks one warp per CTA spreads them over all SMs
    static const int tb_env = std::getenv("LDSR_SCAN_TB") ? std::atoi(std::getenv("LDSR_SCAN_TB")) : 0; // development
    const int TB = tb_env > 0 ? tb_env : ((long long)P.n_fits * P.n_chunks <= 32 * 148 * 2 ? 32 : 64);
    scan_prep_kernel<D><<<(unsigned)((nt + 255) / 256), 256, 0, st>>>(P);
    const int nc = P.n_fits * P.n_chunks;
    const unsigned gb = (unsigned)((nc + TB - 1) / TB);
    // long series: SCAN_GROUPS CTAs per fit scan the chunk elements (three launches), see scan_filt_scan_kernel
    const bool groups = P.n_groups > 1;
    const dim3 gg((unsigned)P.n_fits, (unsigned)P.n_groups);
    if (P.n_chunks > 1) {
        scan_filt_agg_kernel<D><<<gb, TB, 0, st>>>(P);
        if (groups) {
            const size_t smem = (size_t)(SCAN_NT_GROUP + SCAN_NT_GROUP / 32) * FiltElem<D>::LEN * sizeof(double);
            const size_t smem_top = (size_t)SCAN_GROUPS * FiltElem<D>::LEN * sizeof(double);
            LDSR_TRY(opt_in(scan_filt_scan_kernel<D, SCAN_NT_GROUP, 1>, smem));
            LDSR_TRY(opt_in(scan_filt_scan_kernel<D, SCAN_NT_GROUP, 2>, smem));
            scan_filt_scan_kernel<D, SCAN_NT_GROUP, 1><<<gg, SCAN_NT_GROUP, smem, st>>>(P);
            scan_filt_top_kernel<D><<<P.n_fits, SCAN_GROUPS, smem_top, st>>>(P);
            scan_filt_scan_kernel<D, SCAN_NT_GROUP, 2><<<gg, SCAN_NT_GROUP, smem, st>>>(P);
        } else {
            const size_t smem = (size_t)(SCAN_NT + SCAN_NT / 32) * FiltElem<D>::LEN * sizeof(double);
            LDSR_TRY(opt_in(scan_filt_scan_kernel<D, SCAN_NT, 0>, smem));
            scan_filt_scan_kernel<D, SCAN_NT, 0><<<P.n_fits, SCAN_NT, smem, st>>>(P);
        }
    }
    scan_filt_down_kernel<D><<<gb, TB, 0, st>>>(P);
    if (P.n_chunks > 1) {
        scan_smth_agg_kernel<D><<<gb, TB, 0, st>>>(P);
        if (groups) {
            const size_t smem = (size_t)(SCAN_NT_GROUP + SCAN_NT_GROUP / 32) * SmthElem<D>::LEN * sizeof(double);
            const size_t smem_top = (size_t)SCAN_GROUPS * SmthElem<D>::LEN * sizeof(double);
            LDSR_TRY(opt_in(scan_smth_scan_kernel<D, SCAN_NT_GROUP, 1>, smem));
            LDSR_TRY(opt_in(scan_smth_scan_kernel<D, SCAN_NT_GROUP, 2>, smem));
            scan_smth_scan_kernel<D, SCAN_NT_GROUP, 1><<<gg, SCAN_NT_GROUP, smem, st>>>(P);
            scan_smth_top_kernel<D><<<P.n_fits, SCAN_GROUPS, smem_top, st>>>(P);
            scan_smth_scan_kernel<D, SCAN_NT_GROUP, 2><<<gg, SCAN_NT_GROUP, smem, st>>>(P);
        } else {
            const size_t smem = (size_t)(SCAN_NT + SCAN_NT / 32) * SmthElem<D>::LEN * sizeof(double);
            LDSR_TRY(opt_in(scan_smth_scan_kernel<D, SCAN_NT, 0>, smem));
            scan_smth_scan_kernel<D, SCAN_NT, 0><<<P.n_fits, SCAN_NT, smem, st>>>(P);
        }
    }
    scan_smth_down_kernel<D><<<gb, TB, 0, st>>>(P);
    scan_lik_kernel<D><<<P.n_fits, SCAN_LIK_NT, 0, st>>>(P);
    return cudaGetLastError();
}

} // namespace

int scan_groups_for(int n_chunks) { return n_chunks >= SCAN_GROUPS_FROM ? SCAN_GROUPS : 1; }

cudaError_t scan_smoother_launch(int D, const ScanParams &P, cudaStream_t st) {
    switch (D) {
    case 1: return run<1>(P, st);
    case 2: return run<2>(P, st);
    case 3: return run<3>(P, st);
    case 4: return run<4>(P, st);
    default: return cudaErrorInvalidValue;
    }
}

} // namespace ldsr
