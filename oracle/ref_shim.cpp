/*
 * ref_shim.cpp -- read-only window into the UNMODIFIED reference build.
 *
 * Compiled together with the reference translation units (straight from
 * /root/reference/src, never copied) into oracle/_ref/libjincresize_ref.so.
 * It includes the reference's own JincResize.h so tests can look at the tables
 * the reference built for a filter instance: EWAPixelCoeff{meta,factor,
 * filter_size,coeff_stride} (src/JincResize.h:11-25) and the double LUT
 * (src/JincResize.h:27-37).  Test infrastructure only.
 */
#include "JincResize.h"

extern "C" {

__attribute__((visibility("default"))) int ref_table_count(AVS_FilterInfo* fi)
{
    auto* d = reinterpret_cast<JincResize*>(fi->user_data);
    return d ? static_cast<int>(d->out.size()) : 0;
}

__attribute__((visibility("default"))) int ref_table_view(AVS_FilterInfo* fi, int k, int* filter_size, int* coeff_stride,
                                                          const int** meta, const float** factor)
{
    auto* d = reinterpret_cast<JincResize*>(fi->user_data);
    if (!d || k < 0 || k >= static_cast<int>(d->out.size()))
        return -1;
    const EWAPixelCoeff* t = d->out[k];
    *filter_size = t->filter_size;
    *coeff_stride = t->coeff_stride;
    *meta = reinterpret_cast<const int*>(t->meta); /* {start_x, start_y, coeff_meta} per output pixel */
    *factor = t->factor;
    return 0;
}

__attribute__((visibility("default"))) const double* ref_lut(AVS_FilterInfo* fi)
{
    auto* d = reinterpret_cast<JincResize*>(fi->user_data);
    return d ? d->init_lut->lut : nullptr;
}

__attribute__((visibility("default"))) float ref_peak(AVS_FilterInfo* fi)
{
    auto* d = reinterpret_cast<JincResize*>(fi->user_data);
    return d ? d->peak : 0.f;
}

} /* extern "C" */
