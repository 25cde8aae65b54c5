// scan_inst.cu -- instantiations and launch sequence of the general-state-dimension smoother.
#include "scan_kernels.cuh"

namespace ldsr {
namespace {

template <int D> cudaError_t run(const ScanParams &P, cudaStream_t st) {
    const int TB = 64;
    const long long nt = (long long)P.n_fits * P.T;
    scan_prep_kernel<D><<<(unsigned)((nt + 255) / 256), 256, 0, st>>>(P);
    const int nc = P.n_fits * P.n_chunks;
    const unsigned gb = (unsigned)((nc + TB - 1) / TB);
    if (P.n_chunks > 1) {
        scan_filt_agg_kernel<D><<<gb, TB, 0, st>>>(P);
        {
            const size_t smem = (size_t)(SCAN_NT + SCAN_NT / 32) * FiltElem<D>::LEN * sizeof(double);
            cudaError_t e = cudaFuncSetAttribute(scan_filt_scan_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            scan_filt_scan_kernel<D><<<P.n_fits, SCAN_NT, smem, st>>>(P);
        }
    }
    scan_filt_down_kernel<D><<<gb, TB, 0, st>>>(P);
    if (P.n_chunks > 1) {
        scan_smth_agg_kernel<D><<<gb, TB, 0, st>>>(P);
        {
            const size_t smem = (size_t)(SCAN_NT + SCAN_NT / 32) * SmthElem<D>::LEN * sizeof(double);
            cudaError_t e = cudaFuncSetAttribute(scan_smth_scan_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            scan_smth_scan_kernel<D><<<P.n_fits, SCAN_NT, smem, st>>>(P);
        }
    }
    scan_smth_down_kernel<D><<<gb, TB, 0, st>>>(P);
    scan_lik_kernel<D><<<P.n_fits, 32, 0, st>>>(P);
    return cudaGetLastError();
}

} // namespace

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
