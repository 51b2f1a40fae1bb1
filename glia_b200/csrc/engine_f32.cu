// engine_f32.cu -- single-precision instantiation of the engine and its kernels.
#include "engine.cuh"
#include "spectral3d.cuh"
namespace glia {
EngineBase* make_engine_f32(const int n[3], int device, double dt_ctx, int rank, int nranks, int nbatch) {
  return new Engine<float>(n, device, dt_ctx, rank, nranks, nbatch);
}
}  // namespace glia
