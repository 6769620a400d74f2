// zstd_dec_host.cpp — test infrastructure: compiles the DEVICE frame decoder of longtail_b200/csrc/zstd_dec.cu for the host
// (one lane; warp primitives become identities) so the CPU suite can check its serial logic — header / table / bitstream parsing,
// repcode rules, sequence execution — against frames the reference's encoder wrote.  The GPU tests check the same code as a warp.
#include <stdint.h>
#include <string.h>

#define LT_ZSTD_DEC_HOST 1
#define ZD_LANES 1u
#define __device__
#define __forceinline__ inline
#define __constant__ static const
static inline int __clz(uint32_t v) { return v ? __builtin_clz(v) : 32; }
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t s) { return (uint32_t)((((uint64_t)hi << 32) | lo) >> (s & 31u)); }
template <typename T> static inline T __shfl_sync(uint32_t, T v, int) { return v; }
static inline bool __any_sync(uint32_t, bool p) { return p; }
static inline void __syncwarp() {}

#include "../../longtail_b200/csrc/zstd_dec.cu"

extern "C" uint32_t zd_host_worker_bytes() { return (uint32_t)sizeof(ltb::ZstdDecWorker); }

// src must be readable for 8 bytes past n and start 4-byte aligned slack-safe (the caller pads); returns the size or 0xffffffff
extern "C" uint32_t zd_host_decode(void* worker, const uint8_t* src, uint32_t n, uint8_t* dst, uint32_t cap)
{
    return ltb::zstd_decode_frame(static_cast<ltb::ZstdDecWorker*>(worker), src, n, dst, cap, 0);
}
