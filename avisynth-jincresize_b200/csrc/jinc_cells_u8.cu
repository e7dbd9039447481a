// piecewise-periodic rational-ratio kernels for uint8_t planes
#include "jinc_cells.cuh"

namespace jinc_rs {
template int launch_cells<uint8_t>(const jinc_table*, CellsArgs&, int, cudaStream_t, const Rect*, int);
}
