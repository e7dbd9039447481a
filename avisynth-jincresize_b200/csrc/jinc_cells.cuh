// jinc_cells.cuh -- interior kernel for rational ratios P:Q whose phase pattern is piecewise periodic (3:2 = 720p -> 1080p,
// 4:3, 3x, 4x, shifts at 1:1, ...); included by jinc_cells_<type>.cu, which instantiates launch_cells for one sample type.
//
// Replaces JincResize::resize_plane_c (src/JincResize.cpp:560-587) for the pixels whose coefficient block the reference
// finds through factor_map (:431-435, 517-518).  With crop/dst = Q/P output P*c + p of a row reads the window at
// Q*c + off[p] with the phase block of residue p -- exactly so when Q/P is a dyadic fraction (the periodic paths), and
// PIECEWISE so otherwise: the reference accumulates positions in float (:363, 524-528), the quantised phase of a residue
// drifts, and every few dozen cells it steps to the next phase index (1280 -> 1920: 15 phases per axis over 31 runs).
// The table build lays a regular grid of 4-cell groups over each axis and cuts it into CHUNKS inside which every residue
// keeps its phase block and its origins advance by exactly Q per cell: one chunk per group, more where a residue's phase
// steps inside the group (jinc_table.cu, build_cells_axis).  A thread computes all 4 cells of its group and stores the
// ones of its chunk, so the lanes of a warp always read shared memory at the same alignment (the two lanes of a split
// group read the same words: a broadcast, not a bank conflict).
//
// One thread owns one x-chunk x one y-chunk and walks the P x P residue pairs one after the other: the 4 x 4 outputs of
// a pair share ONE weight block, held in registers (FS*FS floats, read once per pass from L1/L2), and their windows
// overlap, so a row of Q*3 + FS source values feeds 4 x FS FMAs of up to 4 output rows.  The source footprint of the
// whole tile (32 x-chunks x WARPS y-chunks, all residues) is staged ONCE into shared memory as floats, columns
// de-interleaved by (c mod 4Q) so that the lanes of a warp -- consecutive chunks, 4Q columns apart -- read consecutive
// words.  The row loop is fully unrolled: every shared-memory address is a per-pass register plus an immediate.
#ifndef JINC_CELLS_CUH
#define JINC_CELLS_CUH

#include <climits>

#include "jinc_resample.cuh"

namespace jinc_rs {

template <int FS, int Q>
struct CellsGeom {
    static constexpr int NX = JINC_CELLS_NX, NY = JINC_CELLS_NY;
    static constexpr int WARPS = jinc_cells_warps(Q, FS);
    static constexpr int THREADS = 32 * WARPS;
    static constexpr int FSP = (FS + 3) & ~3;
    static constexpr int SPAN = Q * (NX - 1) + FS; // columns a thread reads per source row
    static constexpr int NROW = Q * (NY - 1) + FS; // source rows a thread walks per pass
    static constexpr int D = Q * NX;               // column de-interleave modulus
    static constexpr int FW = jinc_cells_footprint(Q, FS, NX, 32);    // staged columns (worst case, checked by the table build)
    static constexpr int FH = jinc_cells_footprint(Q, FS, NY, WARPS); // staged rows
    static constexpr int SUB = (FW + D - 1) / D;
    static constexpr int ROW = D * SUB;
    static constexpr size_t SMEM = (size_t)FH * ROW * sizeof(float);
};


// the NX x NY samples of one residue pair; PX > 0: the cell size is a compile-time constant (immediate store offsets)
template <typename T, int NX, int NY, int PX>
__device__ __forceinline__ void store_cells(T* __restrict__ o, long long rstep, int px_runtime, unsigned live, const float (&acc)[NY][NX],
                                            float peak)
{
    const int step = PX > 0 ? PX : px_runtime;
#pragma unroll
    for (int j = 0; j < NY; ++j, o += rstep)
#pragma unroll
        for (int i = 0; i < NX; ++i)
            if (live & (1u << (j * NX + i)))
                o[i * step] = finish<T>(acc[j][i], peak);
}

template <typename T, int FS, int Q>
__global__ void __launch_bounds__((CellsGeom<FS, Q>::THREADS), (CellsGeom<FS, Q>::THREADS == 128 ? (FS <= 7 && Q <= 2 ? 4 : FS <= 9 ? 3 : 2) : 2)) resample_cells(const __grid_constant__ CellsArgs a)
{
    using G = CellsGeom<FS, Q>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* __restrict__ tile = reinterpret_cast<float*>(smem_raw);

    unsigned role_id;
    if (block_role(blockIdx.x, (unsigned)a.strip_blocks, a.strip_shift, role_id)) {
        if (a.st.plan_patches)
            strip_block_planned<T, FS, G::THREADS, CL_STRIP_SPT, Q>(a.st, a.fr, role_id, tile);
        else
            strip_block_unplanned<T, FS, G::THREADS, CL_STRIP_SPT, 0, 0>(a.st, a.fr, role_id, tile);
        return;
    }
    const int plane = (int)div_by(role_id, a.tiles_per_plane_magic);
    const int tidx = role_id - plane * a.tiles_per_plane;
    const int tile_y = (int)div_by((unsigned)tidx, a.tiles_x_magic), tile_x = tidx - tile_y * a.tiles_x;
    const PlanePtrs& pp = frame_ptrs(a.fr);
    const T* __restrict__ src = static_cast<const T*>(pp.src[plane]);
    T* __restrict__ dst = static_cast<T*>(pp.dst[plane]);
    const int sp = (int)pp.src_pitch[plane];
    const long long dp = pp.dst_pitch[plane];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int Px = a.Px, Py = a.Py;

    // ---- footprint origin of the tile: the smallest origin of its first chunks (origins grow with the chunk index)
    const int cxk0 = tile_x * 32, cyk0 = a.cyk_begin + tile_y * G::WARPS;
    int lo_x = INT_MAX, lo_y = INT_MAX;
    for (int p = 0; p < Px; ++p)
        lo_x = min(lo_x, __ldg(a.cx_org + cxk0 * Px + p));
    for (int p = 0; p < Py; ++p)
        lo_y = min(lo_y, __ldg(a.cy_org + cyk0 * Py + p));

    // ---- stage the footprint once, converted to float; a warp takes rows, a lane the columns lane + 32 q.  Two rows of
    //      loads are in flight before the first store.
    {
        constexpr int CQ = (G::FW + 31) / 32;
        int gx[CQ], so[CQ];
#pragma unroll
        for (int q = 0; q < CQ; ++q) {
            const int c = lane + 32 * q;
            gx[q] = min(max(lo_x + c, 0), a.src_w - 1); // columns outside the plane only feed chunks that do not exist
            so[q] = (c % G::D) * G::SUB + c / G::D;
        }
        for (int r0 = warp; r0 < G::FH; r0 += 2 * G::WARPS) {
            T v[2][CQ];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int r = min(r0 + u * G::WARPS, G::FH - 1);
                const T* __restrict__ srow = src + (long long)min(max(lo_y + r, 0), a.src_h - 1) * sp;
#pragma unroll
                for (int q = 0; q < CQ; ++q)
                    v[u][q] = __ldg(srow + gx[q]);
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int r = r0 + u * G::WARPS;
                if (r < G::FH) {
#pragma unroll
                    for (int q = 0; q < CQ; ++q)
                        if (G::FW % 32 == 0 || lane + 32 * q < G::FW)
                            tile[r * G::ROW + so[q]] = sample_to_float(v[u][q]);
                }
            }
        }
    }
    __syncthreads();

    const int cxk = cxk0 + lane, cyk = cyk0 + warp;
    if (cxk >= a.n_cx || cyk >= a.cyk_end)
        return;
    const int ncx = __ldg(a.cx_n + cxk), ncy = __ldg(a.cy_n + cyk);     // live cells of the chunk ...
    const int ix0 = __ldg(a.cx_i0 + cxk), iy0 = __ldg(a.cy_i0 + cyk);   // ... from this cell of the group on
    const int cell_x = __ldg(a.cx_cell + cxk), cell_y = __ldg(a.cy_cell + cyk); // first cell of the group
    // output addressing, once per thread: residue (py, px) of cell (cell_x + i, cell_y + j) is obase[py * dp + px + j * rstep + i * Px]
    T* __restrict__ const obase = dst + (long long)(a.y0 + Py * cell_y) * dp + (a.x0 + Px * cell_x);
    const long long rstep = (long long)Py * dp;
    unsigned live = 0; // bit j * NX + i: cell (i, j) of the group belongs to this chunk (and to the row band)
#pragma unroll
    for (int j = 0; j < G::NY; ++j)
#pragma unroll
        for (int i = 0; i < G::NX; ++i)
            if (i >= ix0 && i < ix0 + ncx && j >= iy0 && j < iy0 + ncy && cell_y + j >= a.cell_y_begin && cell_y + j < a.cell_y_end)
                live |= 1u << (j * G::NX + i);

    // The passes run px-major (px outside, py inside: the shared-memory column offsets of a pass depend on px only) as one
    // flat, software-pipelined loop: the weight block of the NEXT pass is loaded into the weight registers as soon as this
    // pass's FMAs are done -- they are dead by then -- and before its samples are converted and stored, so the loads' round
    // trip overlaps the store phase instead of opening the next pass.
    float w[FS][FS]; // the pair's weight block, in registers for the whole pass
    int addr[G::SPAN]; // shared-memory word of column ox + k within a footprint row
    int rx = 0, oy = 0;
    auto load_px = [&](int px) {
        const int ox = __ldg(a.cx_org + cxk * Px + px) - lo_x;
        rx = __ldg(a.cx_rank + cxk * Px + px);
        const int m0 = ox % G::D, q0 = ox / G::D;
#pragma unroll
        for (int k = 0; k < G::SPAN; ++k)
            addr[k] = q0 + ((m0 + k) % G::D) * G::SUB + (m0 + k) / G::D;
    };
    auto load_py = [&](int py) {
        oy = __ldg(a.cy_org + cyk * Py + py) - lo_y;
        const int ry = __ldg(a.cy_rank + cyk * Py + py);
        const float* __restrict__ wb = a.wblocks + (size_t)(ry * a.n_rank_x + rx) * (unsigned)(FS * a.wstride);
        if (a.wstride == G::FSP) {
#pragma unroll
            for (int ly = 0; ly < FS; ++ly) {
#pragma unroll
                for (int q4 = 0; q4 < G::FSP / 4; ++q4) {
                    const float4 t = __ldg(reinterpret_cast<const float4*>(wb + ly * G::FSP) + q4);
                    const float tv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (4 * q4 + e < FS)
                            w[ly][4 * q4 + e] = tv[e];
                }
            }
        } else {
#pragma unroll
            for (int ly = 0; ly < FS; ++ly)
#pragma unroll
                for (int lx = 0; lx < FS; ++lx)
                    w[ly][lx] = __ldg(wb + ly * a.wstride + lx);
        }
    };
    int px = 0, py = 0;
    load_px(0);
    load_py(0);
#pragma unroll 1
    for (int pass = Px * Py; pass > 0; --pass) {
        const float* __restrict__ trow = tile + oy * G::ROW; // the pass's first footprint row
        float acc[G::NY][G::NX];
#pragma unroll
        for (int j = 0; j < G::NY; ++j)
#pragma unroll
            for (int i = 0; i < G::NX; ++i)
                acc[j][i] = 0.f;

#pragma unroll
        for (int r = 0; r < G::NROW; ++r) {
            float s[G::SPAN];
#pragma unroll
            for (int k = 0; k < G::SPAN; ++k)
                s[k] = trow[addr[k] + r * G::ROW];
#pragma unroll
            for (int j = 0; j < G::NY; ++j) {
                const int ly = r - Q * j; // weight row of output row j (a constant after unrolling)
                if (ly >= 0 && ly < FS) {
#pragma unroll
                    for (int lx = 0; lx < FS; ++lx)
#pragma unroll
                        for (int i = 0; i < G::NX; ++i)
                            acc[j][i] = fmaf(s[Q * i + lx], w[ly][lx], acc[j][i]);
                }
            }
        }

        // ---- this residue pair's samples of the chunk: every Px-th column of every Py-th row.  One code path for
        //      every lane (a warp holds whole and split groups side by side): a per-thread bit mask says which of
        //      the 4 x 4 samples belong to the chunk, and the usual cell sizes get immediate store offsets.
        T* __restrict__ o = obase + (long long)py * dp + px;
        // the next pass's offsets and weights, before the stores
        if (++py == Py) {
            py = 0;
            ++px;
            if (pass > 1)
                load_px(px);
        }
        if (pass > 1)
            load_py(py);
        switch (Px) {
        case 3: store_cells<T, G::NX, G::NY, 3>(o, rstep, Px, live, acc, a.fr.peak); break;
        case 4: store_cells<T, G::NX, G::NY, 4>(o, rstep, Px, live, acc, a.fr.peak); break;
        default: store_cells<T, G::NX, G::NY, 0>(o, rstep, Px, live, acc, a.fr.peak); break;
        }
    }
}

template <typename T, int FS, int Q>
int launch_cells_cfg(const jinc_table* t, CellsArgs& a, int n_frames, cudaStream_t st, const Rect* rects, int n_rects)
{
    using G = CellsGeom<FS, Q>;
    auto kern = resample_cells<T, FS, Q>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM);
    if (e != cudaSuccess)
        return jinc_fail(JINC_E_CUDA, "cudaFuncSetAttribute(cells smem %zu): %s", G::SMEM, cudaGetErrorString(e));
    long long strip_blocks =
        n_rects > 0 ? set_strip_rects(a.st, rects, n_rects, G::THREADS * CL_STRIP_SPT, CL_STRIP_MAX_PW, G::SMEM) * a.fr.n_planes : 0;
    if (n_rects > 0 && a.want_strip_plan && attach_strip_plan(t, a.st, G::THREADS, CL_STRIP_SPT))
        strip_blocks = (long long)a.st.blocks_per_plane * a.fr.n_planes;
    a.tiles_x = (a.n_cx + 31) / 32;
    a.tiles_per_plane = a.tiles_x * ((a.cyk_end - a.cyk_begin + G::WARPS - 1) / G::WARPS);
    a.tiles_x_magic = div_magic((unsigned)a.tiles_x);
    a.tiles_per_plane_magic = div_magic((unsigned)a.tiles_per_plane);
    if (a.interior_blocks)
        a.interior_blocks = a.tiles_per_plane * a.fr.n_planes;
    if (a.interior_blocks + strip_blocks == 0)
        return 2;
    a.strip_blocks = (int)strip_blocks;
    a.strip_shift = strip_role_shift(a.interior_blocks, strip_blocks);
    dim3 grid((unsigned)(a.interior_blocks + strip_blocks), n_frames, 1);
    kern<<<grid, G::THREADS, G::SMEM, st>>>(a);
    e = cudaGetLastError();
    if (e != cudaSuccess)
        return jinc_fail(JINC_E_CUDA, "resample_cells launch failed: %s", cudaGetErrorString(e));
    (void)t;
    return JINC_OK;
}

// One source step Q per translation unit (jinc_cells_<type>_q<Q>.cu), so the unrolled kernels build in parallel.
// 0 launched, 2 nothing to do, 1 unsupported geometry, <0 error
template <typename T, int Q>
int launch_cells_q(const jinc_table* t, CellsArgs& a, int n_frames, cudaStream_t st, const Rect* rects, int n_rects)
{
    switch (t->sc.fs) {
    case 5: // tap 2
        if constexpr (jinc_cells_instantiated(Q, 5))
            return launch_cells_cfg<T, 5, Q>(t, a, n_frames, st, rects, n_rects);
        break;
    case 7: // tap 3 at upscale ratios
        if constexpr (jinc_cells_instantiated(Q, 7))
            return launch_cells_cfg<T, 7, Q>(t, a, n_frames, st, rects, n_rects);
        break;
    case 9: // tap 4 at upscale ratios; tap 3 at 3:4
        if constexpr (jinc_cells_instantiated(Q, 9))
            return launch_cells_cfg<T, 9, Q>(t, a, n_frames, st, rects, n_rects);
        break;
    case 10: // tap 3 at 2:3
        if constexpr (jinc_cells_instantiated(Q, 10))
            return launch_cells_cfg<T, 10, Q>(t, a, n_frames, st, rects, n_rects);
        break;
    case 11: // tap 5
        if constexpr (jinc_cells_instantiated(Q, 11))
            return launch_cells_cfg<T, 11, Q>(t, a, n_frames, st, rects, n_rects);
        break;
    default: break;
    }
    return 1;
}

} // namespace jinc_rs

#endif
