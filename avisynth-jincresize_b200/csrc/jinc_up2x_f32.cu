// exact-2x interior kernels for float planes
#include "jinc_up2x.cuh"

namespace jinc_rs {
template int launch_up2x<float>(const jinc_table*, UpArgs&, long long, int, cudaStream_t);
}
