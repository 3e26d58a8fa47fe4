// rt_kernels_f64.cu — double instantiations.  This translation unit is compiled with --fmad=false:
// every product and sum rounds separately, in the order the reference's Python evaluates them, so
// hit decisions of the deterministic renderers are the reference's own (bit-exact hit index).
#include "rt_kernels.cuh"
#include "rt_resolve_hybrid.cuh"

template <> cudaError_t launch_resolve<double>(const SceneView<double>& sc, const RenderArgs& a, cudaStream_t st, LaunchInfo* info) {
  return launch_resolve_generic<double>(sc, a, st, info);
}
template cudaError_t launch_pt_mega<double>(const SceneView<double>&, const RenderArgs&, cudaStream_t, LaunchInfo*);
template cudaError_t launch_probe<double>(const SceneView<double>&, const RenderArgs&, const ProbeArgs&, cudaStream_t);

size_t resolve_hybrid_table_bytes(int n_spheres, int n_lights) {
  return (size_t)(1 + n_lights) * (size_t)((n_spheres + 1) / 2) * RT_CO_REC_BYTES + 256;
}
cudaError_t launch_resolve_hybrid(const SceneView<double>& sc, const RenderArgs& a, float* co, cudaStream_t st, LaunchInfo* info) {
  return launch_resolve_hybrid_impl(sc, a, co, st, info);
}

// FP64 FMA throughput probe: the denominator for the fp64 / hybrid kernels' fp64 share.  Explicit __fma_rn,
// so it is DFMA even in this translation unit (the renderers themselves never fuse: every product and sum
// rounds separately, i.e. they can reach at most half of this figure in flops).
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1.0, x2 = x0 + 2.0, x3 = x0 + 3.0, x4 = x0 + 4.0, x5 = x0 + 5.0, x6 = x0 + 6.0, x7 = x0 + 7.0;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      x0 = __fma_rn(x0, a, b); x1 = __fma_rn(x1, a, b); x2 = __fma_rn(x2, a, b); x3 = __fma_rn(x3, a, b);
      x4 = __fma_rn(x4, a, b); x5 = __fma_rn(x5, a, b); x6 = __fma_rn(x6, a, b); x7 = __fma_rn(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

cudaError_t launch_dfma(double* out, int blocks, int iters, cudaStream_t st) {
  k_dfma<<<blocks, 256, 0, st>>>(out, iters, 0.999, 0.001);
  return cudaGetLastError();
}
