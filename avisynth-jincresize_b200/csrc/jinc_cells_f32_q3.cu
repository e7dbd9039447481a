// piecewise-periodic rational-ratio kernels, source step 3 per cell, for float planes
#include "jinc_cells.cuh"

namespace jinc_rs {
template int launch_cells_q<float, 3>(const jinc_table*, CellsArgs&, int, cudaStream_t, const Rect*, int);
}
