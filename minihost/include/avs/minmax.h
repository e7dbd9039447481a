/* avs/minmax.h -- clean-room stand-in for the AviSynth+ SDK helper header.
 * Three function templates; clamp tests the upper bound first and lets NaN
 * pass through. */
#ifndef MINIHOST_AVS_MINMAX_H
#define MINIHOST_AVS_MINMAX_H

template <typename T>
T min(T a, T b)
{
    return a < b ? a : b;
}

template <typename T>
T max(T a, T b)
{
    return a > b ? a : b;
}

template <typename T>
T clamp(T n, T lo, T hi)
{
    n = n > hi ? hi : n;
    return n < lo ? lo : n;
}

#endif
