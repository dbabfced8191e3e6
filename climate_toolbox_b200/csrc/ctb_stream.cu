// Entry point of the streaming kernel (ctb_stream_impl.cuh): one translation unit per input type.
#include "ctb_internal.cuh"

int ctb_launch_stream_f32(const ctb_plan* P, const AggArgs& a, int kind, int n_out, cudaStream_t st);
int ctb_launch_stream_f64(const ctb_plan* P, const AggArgs& a, int kind, int n_out, cudaStream_t st);

int ctb_launch_stream(const ctb_plan* P, const AggArgs& a, int dtype, int kind, int n_out, cudaStream_t st) {
  return dtype == CTB_F32 ? ctb_launch_stream_f32(P, a, kind, n_out, st) : ctb_launch_stream_f64(P, a, kind, n_out, st);
}
