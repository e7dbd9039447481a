// piecewise-periodic rational-ratio kernels for float planes
#include "jinc_cells.cuh"

namespace jinc_rs {
template int launch_cells<float>(const jinc_table*, CellsArgs&, int, cudaStream_t, const Rect*, int);
}
