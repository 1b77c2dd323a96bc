// shim pch.h for building the reference's physics sources with g++ (replaces src/pch.h: Windows / WRL headers removed)
#pragma once
#include <cstdint>
#include <cstddef>
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <math.h>      // libstdc++'s wrapper: puts the float overloads (sqrt, abs, sin, ...) into the GLOBAL namespace like MSVC's <cmath> does;
#include <stdlib.h>    // without it `sqrt(float)` / `abs(float)` in the reference would silently pick the C double / int functions
#include <cfloat>
#include <climits>
#include <limits>
#include <array>
#include <string>
#include <vector>
#include <iostream>
#include <memory>
#include <mutex>
#include <functional>
#include <algorithm>
#include <type_traits>
#include <utility>
#include <filesystem>
namespace fs = std::filesystem;

typedef int8_t int8;
typedef uint8_t uint8;
typedef int16_t int16;
typedef uint16_t uint16;
typedef int32_t int32;
typedef uint32_t uint32;
typedef int64_t int64;
typedef uint64_t uint64;
typedef wchar_t wchar;

extern "C" void refAssertFailed(const char* cond, const char* file, int line);
#define ASSERT(cond) (void)((!!(cond)) || (refAssertFailed(#cond, __FILE__, __LINE__), 0))
#define __debugbreak() ((void)0)
#define __declspec(x)

template <typename T> using ref = std::shared_ptr<T>;
template <typename T> using weakref = std::weak_ptr<T>;
template <typename T, typename... Args> inline ref<T> make_ref(Args&&... args) { return std::make_shared<T>(std::forward<Args>(args)...); }

// arraysize must also work on the array views that replace the reference's "named members / array" unions (ref_array_view below)
template <typename T> struct ref_array_count;
template <typename T, size_t N> struct ref_array_count<T[N]> { static constexpr size_t value = N; };
#define arraysize(arr) (ref_array_count<std::remove_cv_t<std::remove_reference_t<decltype(arr)>>>::value)

template <typename T> constexpr inline auto min(T a, T b) { return (a < b) ? a : b; }
template <typename T> constexpr inline auto max(T a, T b) { return (a < b) ? b : a; }
template <auto V> static constexpr auto force_consteval = V;

#define setBit(mask, bit) (mask) |= (1 << (bit))
#define unsetBit(mask, bit) (mask) ^= (1 << (bit))

// SVML integer division intrinsics (MSVC ships them; g++ does not): lane-wise scalar division, same results
#include <immintrin.h>
static inline __m128i _mm_div_epi32(__m128i a, __m128i b) { alignas(16) int x[4], y[4]; _mm_store_si128((__m128i*)x, a); _mm_store_si128((__m128i*)y, b); for (int i = 0; i < 4; ++i) x[i] /= y[i]; return _mm_load_si128((__m128i*)x); }
#ifdef __AVX2__
static inline __m256i _mm256_div_epi32(__m256i a, __m256i b) { alignas(32) int x[8], y[8]; _mm256_store_si256((__m256i*)x, a); _mm256_store_si256((__m256i*)y, b); for (int i = 0; i < 8; ++i) x[i] /= y[i]; return _mm256_load_si256((__m256i*)x); }
#endif

// indexable view over the N members of type T declared immediately BEFORE it in the same struct (see patch_reference.py,
// union_to_struct): stands in for the array arm of the reference's "named members / array" unions
template <typename T, int N> struct ref_array_view
{
    T& operator[](size_t i) { return (reinterpret_cast<T*>(this) - N)[i]; }
    const T& operator[](size_t i) const { return (reinterpret_cast<const T*>(this) - N)[i]; }
};
template <typename T, int N> struct ref_array_count<ref_array_view<T, N>> { static constexpr size_t value = N; };
// same idea, placed FIRST in a struct with [[no_unique_address]] so that it adds no size: view over the members that follow it
template <typename T> struct ref_array_front
{
    T& operator[](size_t i) { return reinterpret_cast<T*>(this)[i]; }
    const T& operator[](size_t i) const { return reinterpret_cast<const T*>(this)[i]; }
};
// MSVC exposes every intrinsic regardless of /arch; g++ ties the *_epi32 unaligned load/store names to AVX-512VL. Same operation:
#define _mm256_loadu_epi32(p) _mm256_loadu_si256((const __m256i*)(p))
#define _mm256_storeu_epi32(p, v) _mm256_storeu_si256((__m256i*)(p), (v))
#define _mm_loadu_epi32(p) _mm_loadu_si128((const __m128i*)(p))
#define _mm_storeu_epi32(p, v) _mm_storeu_si128((__m128i*)(p), (v))
