// M31 / CM31 / QM31 arithmetic for device and host code.
//
// Field definitions follow upstream stwo core/fields/{m31,cm31,qm31}.rs (un-vendored dependency of the
// reference, pinned in /root/reference/stwo/Cargo.toml:16): p = 2^31-1, CM31 = M31[i]/(i^2+1),
// QM31 = CM31[u]/(u^2-(2+i)).  Canonical representatives in [0,p) everywhere a value is stored.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define HD __host__ __device__ __forceinline__
#else
#define HD inline
#endif

namespace m31 {

constexpr uint32_t P = 0x7fffffffu;

// [0, 2^32) -> [0, p]; composing twice gives a value that umin-fixes to [0,p)
HD uint32_t fold32(uint32_t x) { return (x & P) + (x >> 31); }

HD uint32_t reduce_p(uint32_t x) {  // x in [0, 2p) -> [0,p)
    uint32_t y = x - P;
    return y < x ? y : x;  // unsigned min: if x >= p then y = x-p < x, else y wraps huge
}

HD uint32_t add(uint32_t a, uint32_t b) { return reduce_p(a + b); }
HD uint32_t sub(uint32_t a, uint32_t b) {
    uint32_t d = a - b;
    uint32_t e = d + P;
    return e < d ? e : d;  // if a<b, d wrapped (huge) and e = d+p wraps back small
}
HD uint32_t neg(uint32_t a) { return a == 0 ? 0 : P - a; }

HD uint32_t reduce64(uint64_t v) {  // v < 2^62 -> [0,p)
    uint32_t lo = (uint32_t)v & P;
    uint32_t hi = (uint32_t)(v >> 31);
    return reduce_p(lo + hi);  // lo + hi < 2^32, < 2p since hi <= p-1... (v < p^2 => hi < p)
}

HD uint32_t mul(uint32_t a, uint32_t b) { return reduce64((uint64_t)a * b); }

// general 64-bit value -> [0,p)
HD uint32_t reduce64_full(uint64_t v) {
    uint64_t t = (v & P) + (v >> 31);            // < 2^34
    uint32_t u = (uint32_t)(t & P) + (uint32_t)(t >> 31);  // < 2^31 + 8
    return reduce_p(u);
}

HD uint32_t pow(uint32_t a, uint32_t e) {
    uint32_t r = 1;
    while (e) {
        if (e & 1) r = mul(r, a);
        a = mul(a, a);
        e >>= 1;
    }
    return r;
}
HD uint32_t inv(uint32_t a) { return pow(a, P - 2); }

struct CM31 {
    uint32_t a, b;
};
HD CM31 cadd(CM31 x, CM31 y) { return {add(x.a, y.a), add(x.b, y.b)}; }
HD CM31 csub(CM31 x, CM31 y) { return {sub(x.a, y.a), sub(x.b, y.b)}; }
HD CM31 cmul(CM31 x, CM31 y) {
    // (a+bi)(c+di) = (ac - bd) + (ad + bc)i, lazily reduced through 64-bit sums (each product < 2^62)
    uint64_t ac = (uint64_t)x.a * y.a, bd = (uint64_t)x.b * y.b;
    uint64_t ad = (uint64_t)x.a * y.b, bc = (uint64_t)x.b * y.a;
    return {sub(reduce64(ac), reduce64(bd)), reduce64_full(ad + bc)};
}
HD CM31 cmul_m(CM31 x, uint32_t m) { return {mul(x.a, m), mul(x.b, m)}; }
HD CM31 cneg(CM31 x) { return {neg(x.a), neg(x.b)}; }
HD CM31 cinv(CM31 x) {
    uint32_t n = inv(add(mul(x.a, x.a), mul(x.b, x.b)));
    return {mul(x.a, n), mul(neg(x.b), n)};
}

struct QM31 {
    uint32_t v[4];
};
HD QM31 qzero() { return {{0, 0, 0, 0}}; }
HD QM31 qone() { return {{1, 0, 0, 0}}; }
HD QM31 qfrom(uint32_t a) { return {{a, 0, 0, 0}}; }
HD QM31 qadd(QM31 x, QM31 y) { return {{add(x.v[0], y.v[0]), add(x.v[1], y.v[1]), add(x.v[2], y.v[2]), add(x.v[3], y.v[3])}}; }
HD QM31 qsub(QM31 x, QM31 y) { return {{sub(x.v[0], y.v[0]), sub(x.v[1], y.v[1]), sub(x.v[2], y.v[2]), sub(x.v[3], y.v[3])}}; }
HD QM31 qneg(QM31 x) { return {{neg(x.v[0]), neg(x.v[1]), neg(x.v[2]), neg(x.v[3])}}; }
HD QM31 qmul_m(QM31 x, uint32_t m) { return {{mul(x.v[0], m), mul(x.v[1], m), mul(x.v[2], m), mul(x.v[3], m)}}; }
HD QM31 qmul_c(QM31 x, CM31 c) {
    CM31 lo = cmul({x.v[0], x.v[1]}, c), hi = cmul({x.v[2], x.v[3]}, c);
    return {{lo.a, lo.b, hi.a, hi.b}};
}
HD QM31 qmul(QM31 x, QM31 y) {
    CM31 a0{x.v[0], x.v[1]}, a1{x.v[2], x.v[3]}, b0{y.v[0], y.v[1]}, b1{y.v[2], y.v[3]};
    CM31 t = cmul(a1, b1);
    // (2+i) * t
    CM31 rt{sub(add(t.a, t.a), t.b), add(add(t.b, t.b), t.a)};
    CM31 lo = cadd(cmul(a0, b0), rt);
    CM31 hi = cadd(cmul(a0, b1), cmul(a1, b0));
    return {{lo.a, lo.b, hi.a, hi.b}};
}
HD QM31 qinv(QM31 x) {
    CM31 a{x.v[0], x.v[1]}, b{x.v[2], x.v[3]};
    CM31 b2 = cmul(b, b);
    CM31 rb2{sub(add(b2.a, b2.a), b2.b), add(add(b2.b, b2.b), b2.a)};
    CM31 den = csub(cmul(a, a), rb2);
    CM31 di = cinv(den);
    CM31 lo = cmul(a, di), hi = cmul(cneg(b), di);
    return {{lo.a, lo.b, hi.a, hi.b}};
}
HD QM31 qconj(QM31 x) { return {{x.v[0], x.v[1], neg(x.v[2]), neg(x.v[3])}}; }
HD bool qeq(QM31 x, QM31 y) { return x.v[0] == y.v[0] && x.v[1] == y.v[1] && x.v[2] == y.v[2] && x.v[3] == y.v[3]; }
HD QM31 qpow(QM31 a, uint64_t e) {
    QM31 r = qone();
    while (e) {
        if (e & 1) r = qmul(r, a);
        a = qmul(a, a);
        e >>= 1;
    }
    return r;
}

}  // namespace m31
