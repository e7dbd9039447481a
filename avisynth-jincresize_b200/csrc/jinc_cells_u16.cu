// piecewise-periodic rational-ratio kernels for uint16_t planes
#include "jinc_cells.cuh"

namespace jinc_rs {
template int launch_cells<uint16_t>(const jinc_table*, CellsArgs&, int, cudaStream_t, const Rect*, int);
}
