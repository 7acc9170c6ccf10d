// uint128.h (drop-in) -- source-compatible replacement of the reference's BFV_Scheme/uint128.h for HOST code.
// Same public type (uint128_t with .low/.high), same free functions the reference drivers use (host64x2, exp2,
// operator/ and operator% by a 64-bit value, shifts, comparisons), implemented on the compiler's unsigned __int128
// instead of bit-serial loops.  Device-side 128-bit arithmetic (mul64 / sub128) is gone: kernels live in libnttb200.so.
// Preserved quirks that reference drivers may depend on: `x % y` returns x when x == y, because the
// reference's operator<(uint128_t, uint64_t) is really "<="; operator<<= shifts RIGHT.
#pragma once
#include <cinttypes>
#include <cmath>
#include <cstdint>
#include <string>

class uint128_t {
public:
    unsigned long long low;
    unsigned long long high;

    uint128_t() : low(0), high(0) {}
    uint128_t(const uint64_t &x) : low(x), high(0) {}
    static uint128_t from(unsigned __int128 v) { uint128_t z; z.low = (unsigned long long)v; z.high = (unsigned long long)(v >> 64); return z; }
    unsigned __int128 value() const { return ((unsigned __int128)high << 64) | low; }

    void operator=(const uint128_t &r) { low = r.low; high = r.high; }
    void operator=(const uint64_t &r) { low = r; high = 0; }
    uint128_t operator<<(const unsigned &s) const { return s >= 128 ? uint128_t() : from(value() << s); }
    uint128_t operator>>(const unsigned &s) const { return s >= 128 ? uint128_t() : from(value() >> s); }
    static void shiftr(uint128_t &x, const unsigned &s) { x = x >> s; }
    static uint128_t exp2(const int &e) { return from((unsigned __int128)1 << e); }
    static int log_2(const uint128_t &x) { return x.high ? (int)(std::log2((double)x.high) + 64) : (int)std::log2((double)x.low); }
    static int clz(uint128_t x)
    {
        if (x.high) return __builtin_clzll(x.high);
        return x.low ? 64 + __builtin_clzll(x.low) : 128;
    }
};

static inline void operator<<=(uint128_t &x, const unsigned &s) { x = x >> s; }   // sic: the reference shifts right
static inline bool operator==(const uint128_t &l, const uint128_t &r) { return l.low == r.low && l.high == r.high; }
static inline bool operator<(const uint128_t &l, const uint128_t &r) { return l.value() < r.value(); }
static inline bool operator<(const uint128_t &l, const uint64_t &r) { return l.high == 0 && l.low <= r; }   // sic: "<="
static inline bool operator>(const uint128_t &l, const uint128_t &r) { return l.value() > r.value(); }
static inline bool operator<=(const uint128_t &l, const uint128_t &r) { return l.value() <= r.value(); }
static inline bool operator>=(const uint128_t &l, const uint128_t &r) { return l.value() >= r.value(); }
static inline uint128_t operator+(const uint128_t &x, const uint128_t &y) { return uint128_t::from(x.value() + y.value()); }
static inline uint128_t operator+(const uint128_t &x, const uint64_t &y) { return uint128_t::from(x.value() + y); }
static inline uint128_t operator-(const uint128_t &x, const uint128_t &y) { return uint128_t::from(x.value() - y.value()); }
static inline void operator-=(uint128_t &x, const uint128_t &y) { x = x - y; }
static inline uint128_t operator-(const uint128_t &x, const uint64_t &y) { return uint128_t::from(x.value() - y); }
static inline uint128_t operator/(uint128_t x, const uint64_t &y) { return uint128_t::from(x.value() / y); }
static inline uint128_t operator%(uint128_t x, const uint64_t &y)
{
    if (x < y) return x;                       // includes x == y (reference behaviour)
    return uint128_t::from(x.value() % y);
}
// 64 x 64 -> 128 product (the reference's shift-and-add host64x2)
static inline uint128_t host64x2(const uint64_t &x, const uint64_t &y) { return uint128_t::from((unsigned __int128)x * y); }
