// integer-ratio downscale kernels for uint16_t planes
#include "jinc_down.cuh"

namespace jinc_rs {
template int launch_down<uint16_t>(const jinc_table*, DownArgs&, int, const int*, bool, int, cudaStream_t, const Rect*, int);
}
