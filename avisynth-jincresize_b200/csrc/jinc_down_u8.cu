// integer-ratio downscale kernels for uint8_t planes
#include "jinc_down.cuh"

namespace jinc_rs {
template int launch_down<uint8_t>(const jinc_table*, DownArgs&, int, const int*, bool, int, cudaStream_t, const Rect*, int);
}
