// exact-2x interior kernels for uint16_t planes
#include "jinc_up2x.cuh"

namespace jinc_rs {
template int launch_up2x<uint16_t>(const jinc_table*, UpArgs&, long long, int, cudaStream_t);
}
