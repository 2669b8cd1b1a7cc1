// st_common.cuh -- device helpers shared by the kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "st_internal.h"

#define ST_DISPATCH(DTYPE, ...)                                                      \
    switch (DTYPE) {                                                                 \
        case ST_I16: { using T = int16_t; __VA_ARGS__; } break;                     \
        case ST_I32: { using T = int32_t; __VA_ARGS__; } break;                     \
        case ST_I64: { using T = int64_t; __VA_ARGS__; } break;                     \
        case ST_U8:  { using T = uint8_t; __VA_ARGS__; } break;                     \
        case ST_U16: { using T = uint16_t; __VA_ARGS__; } break;                    \
        case ST_U32: { using T = uint32_t; __VA_ARGS__; } break;                    \
        case ST_U64: { using T = uint64_t; __VA_ARGS__; } break;                    \
        case ST_F32: { using T = float; __VA_ARGS__; } break;                       \
        case ST_F64: { using T = double; __VA_ARGS__; } break;                      \
        default: return cudaErrorInvalidValue;                                       \
    }

namespace st {

template <typename T> struct is_float_t { static constexpr bool value = false; };
template <> struct is_float_t<float> { static constexpr bool value = true; };
template <> struct is_float_t<double> { static constexpr bool value = true; };

template <typename T> struct is_signed_t { static constexpr bool value = T(-1) < T(0); };

// NumPy's default accumulator for ndarray.sum (sumtree_array.py:514 fallback)
template <typename T, bool F = is_float_t<T>::value, bool S = is_signed_t<T>::value> struct fold_acc;
template <typename T, bool S> struct fold_acc<T, true, S> { using type = T; };
template <typename T> struct fold_acc<T, false, true> { using type = long long; };
template <typename T> struct fold_acc<T, false, false> { using type = unsigned long long; };

// level of heap node h (root = level 0)
__host__ __device__ __forceinline__ int heap_level(uint32_t h)
{
#ifdef __CUDA_ARCH__
    return 31 - __clz((int)h);
#else
    return 31 - __builtin_clz(h);
#endif
}

// Arithmetic in the element type, exactly as the reference's C does it: operands promoted
// by the usual C rules, result converted back to T (wraps for the narrow integer types).
// Floats: one IEEE add/sub, never contracted (the file is compiled with -fmad=false too).
template <typename T> __device__ __forceinline__ T t_add(T a, T b)
{
    if constexpr (sizeof(T) == 4 && is_float_t<T>::value) return __fadd_rn(a, b);
    else if constexpr (sizeof(T) == 8 && is_float_t<T>::value) return __dadd_rn(a, b);
    else return (T)(a + b);
}
template <typename T> __device__ __forceinline__ T t_sub(T a, T b)
{
    if constexpr (sizeof(T) == 4 && is_float_t<T>::value) return __fsub_rn(a, b);
    else if constexpr (sizeof(T) == 8 && is_float_t<T>::value) return __dsub_rn(a, b);
    else return (T)(a - b);
}

static inline unsigned int grid_for(int64_t work, int block, int64_t cap)
{
    int64_t g = (work + block - 1) / block;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (unsigned int)g;
}

}  // namespace st
