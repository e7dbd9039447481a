// jinc_border.h -- slots of the per-border-pixel normalisers (shared by host and device code).
#ifndef JINC_BORDER_H
#define JINC_BORDER_H

#ifdef __CUDACC__
#define JINC_HD __host__ __device__
#else
#define JINC_HD
#endif

// ---------------------------------------------------------------------------------------------------------------
// Border pixels (window clamped on either axis) are the only ones without a shared phase block.  Their normaliser
// (the reference's `divider`, a float running sum over the window in row-major order, :439,493) is computed once per
// table and kept per pixel; this maps a border pixel to its slot.  [bx0,bx1) x [by0,by1) is the non-border core.
struct BorderGeom {
    int W, H;
    int bx0, bx1, by0, by1;
    long long off_bottom, off_left, off_right, total; // slot offsets of the four strips
};

JINC_HD inline long long jinc_border_slot(const BorderGeom& g, int x, int y)
{
    if (y < g.by0)
        return (long long)y * g.W + x;
    if (y >= g.by1)
        return g.off_bottom + (long long)(y - g.by1) * g.W + x;
    if (x < g.bx0)
        return g.off_left + (long long)(y - g.by0) * g.bx0 + x;
    return g.off_right + (long long)(y - g.by0) * (g.W - g.bx1) + (x - g.bx1);
}


#endif
