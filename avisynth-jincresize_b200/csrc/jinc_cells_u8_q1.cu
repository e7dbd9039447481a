// piecewise-periodic rational-ratio kernels, source step 1 per cell, for uint8_t planes
#include "jinc_cells.cuh"

namespace jinc_rs {
template int launch_cells_q<uint8_t, 1>(const jinc_table*, CellsArgs&, int, cudaStream_t, const Rect*, int);
}
