// integer-ratio downscale kernels for float planes
#include "jinc_down.cuh"

namespace jinc_rs {
template int launch_down<float>(const jinc_table*, DownArgs&, int, const int*, bool, int, cudaStream_t, const Rect*, int);
}
