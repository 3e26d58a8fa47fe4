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
