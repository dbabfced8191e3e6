// The streaming kernel's double instantiations (see ctb_stream_impl.cuh).
#define CTB_STREAM_TIN double
#define CTB_STREAM_ENTRY ctb_launch_stream_f64
#include "ctb_stream_impl.cuh"
