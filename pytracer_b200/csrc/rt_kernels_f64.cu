// rt_kernels_f64.cu — double instantiations.  This translation unit is compiled with --fmad=false:
// every product and sum rounds separately, in the order the reference's Python evaluates them, so
// hit decisions of the deterministic renderers are the reference's own (bit-exact hit index).
#include "rt_kernels.cuh"

template <> cudaError_t launch_resolve<double>(const SceneView<double>& sc, const RenderArgs& a, cudaStream_t st, LaunchInfo* info) {
  return launch_resolve_generic<double>(sc, a, st, info);
}
template cudaError_t launch_pt_mega<double>(const SceneView<double>&, const RenderArgs&, cudaStream_t, LaunchInfo*);
template cudaError_t launch_probe<double>(const SceneView<double>&, const RenderArgs&, const ProbeArgs&, cudaStream_t);
