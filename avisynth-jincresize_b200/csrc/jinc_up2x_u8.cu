// exact-2x interior kernels for uint8_t planes
#include "jinc_up2x.cuh"

namespace jinc_rs {
template int launch_up2x<uint8_t>(const jinc_table*, UpArgs&, long long, int, cudaStream_t);
}
