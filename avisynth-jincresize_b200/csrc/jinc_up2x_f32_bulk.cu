// exact-2x interior kernels for float planes, tiles staged by the bulk-copy engine (JINCRESIZE_B200_TMA=1)
#include "jinc_up2x.cuh"

namespace jinc_rs {
int launch_up2x_bulkcopy(const jinc_table* t, UpArgs& a, long long strip_blocks, int n_frames, cudaStream_t st)
{
    return launch_up2x_any<float, true>(t, a, strip_blocks, n_frames, st);
}
}
