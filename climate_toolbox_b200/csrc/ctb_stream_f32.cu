// The streaming kernel's float instantiations (see ctb_stream_impl.cuh).
#define CTB_STREAM_TIN float
#define CTB_STREAM_ENTRY ctb_launch_stream_f32
#include "ctb_stream_impl.cuh"
