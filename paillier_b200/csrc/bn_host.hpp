// Minimal unsigned big integer for host-side key setup (context creation).
// Not on the hot path: it derives the per-key constants the kernels need
// (n^2, n^3, Montgomery R, R^2, R^3, -n^-1 mod 2^32, CRT constants, window
// schedules).  The reference gets the same values from gmp.Int
// (/root/reference/paillier.go:72-90, :143-152, thresholdkey.go:63-72).
#pragma once
#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace pgpu {

class BigU {
public:
    std::vector<uint32_t> v;  // little-endian limbs, no trailing zeros

    BigU() {}
    BigU(uint64_t x) { while (x) { v.push_back((uint32_t)x); x >>= 32; } }

    static BigU from_be(const uint8_t* p, size_t len) {
        BigU r;
        r.v.assign((len + 3) / 4, 0);
        for (size_t i = 0; i < len; ++i) {
            size_t pos = len - 1 - i;  // byte significance
            r.v[pos / 4] |= (uint32_t)p[i] << (8 * (pos % 4));
        }
        r.trim();
        return r;
    }
    static BigU from_limbs(const uint32_t* p, size_t n) {
        BigU r; r.v.assign(p, p + n); r.trim(); return r;
    }
    // fixed-width little-endian limb record; throws if it does not fit
    std::vector<uint32_t> limbs(size_t width) const {
        if (v.size() > width) throw std::runtime_error("BigU::limbs: value wider than record");
        std::vector<uint32_t> r(v);
        r.resize(width, 0);
        return r;
    }
    void trim() { while (!v.empty() && v.back() == 0) v.pop_back(); }
    bool is_zero() const { return v.empty(); }
    bool is_odd() const { return !v.empty() && (v[0] & 1); }
    size_t bitlen() const {
        if (v.empty()) return 0;
        return 32 * (v.size() - 1) + (32 - __builtin_clz(v.back()));
    }
    bool bit(size_t i) const { return (i / 32) < v.size() && ((v[i / 32] >> (i % 32)) & 1); }
    uint64_t low64() const { return (v.size() > 0 ? v[0] : 0) | ((uint64_t)(v.size() > 1 ? v[1] : 0) << 32); }

    static int cmp(const BigU& a, const BigU& b) {
        if (a.v.size() != b.v.size()) return a.v.size() < b.v.size() ? -1 : 1;
        for (size_t i = a.v.size(); i-- > 0;)
            if (a.v[i] != b.v[i]) return a.v[i] < b.v[i] ? -1 : 1;
        return 0;
    }
    bool operator<(const BigU& o) const { return cmp(*this, o) < 0; }
    bool operator==(const BigU& o) const { return cmp(*this, o) == 0; }
    bool operator!=(const BigU& o) const { return cmp(*this, o) != 0; }
    bool operator>=(const BigU& o) const { return cmp(*this, o) >= 0; }

    BigU operator+(const BigU& o) const {
        BigU r; size_t n = std::max(v.size(), o.v.size());
        r.v.resize(n + 1);
        uint64_t c = 0;
        for (size_t i = 0; i < n; ++i) {
            c += (uint64_t)(i < v.size() ? v[i] : 0) + (i < o.v.size() ? o.v[i] : 0);
            r.v[i] = (uint32_t)c; c >>= 32;
        }
        r.v[n] = (uint32_t)c; r.trim(); return r;
    }
    // requires *this >= o
    BigU operator-(const BigU& o) const {
        if (*this < o) throw std::runtime_error("BigU: negative result");
        BigU r; r.v.resize(v.size());
        int64_t b = 0;
        for (size_t i = 0; i < v.size(); ++i) {
            int64_t d = (int64_t)v[i] - (i < o.v.size() ? o.v[i] : 0) - b;
            b = d < 0; r.v[i] = (uint32_t)d;
        }
        r.trim(); return r;
    }
    BigU operator*(const BigU& o) const {
        BigU r;
        if (v.empty() || o.v.empty()) return r;
        r.v.assign(v.size() + o.v.size(), 0);
        for (size_t i = 0; i < v.size(); ++i) {
            uint64_t c = 0;
            for (size_t j = 0; j < o.v.size(); ++j) {
                c += (uint64_t)v[i] * o.v[j] + r.v[i + j];
                r.v[i + j] = (uint32_t)c; c >>= 32;
            }
            r.v[i + o.v.size()] = (uint32_t)c;
        }
        r.trim(); return r;
    }
    BigU shl(size_t s) const {
        if (v.empty()) return *this;
        BigU r; size_t w = s / 32, b = s % 32;
        r.v.assign(v.size() + w + 1, 0);
        for (size_t i = 0; i < v.size(); ++i) {
            r.v[i + w] |= v[i] << b;
            if (b) r.v[i + w + 1] |= v[i] >> (32 - b);
        }
        r.trim(); return r;
    }
    BigU shr(size_t s) const {
        BigU r; size_t w = s / 32, b = s % 32;
        if (w >= v.size()) return r;
        r.v.assign(v.size() - w, 0);
        for (size_t i = w; i < v.size(); ++i) {
            r.v[i - w] = v[i] >> b;
            if (b && i + 1 < v.size()) r.v[i - w] |= v[i + 1] << (32 - b);
        }
        r.trim(); return r;
    }
    static BigU pow2(size_t k) { return BigU(1).shl(k); }

    // Knuth algorithm D.  q = a / b, r = a % b.
    static void divmod(const BigU& a, const BigU& b, BigU& q, BigU& r) {
        if (b.v.empty()) throw std::runtime_error("BigU: division by zero");
        if (a < b) { q = BigU(); r = a; return; }
        if (b.v.size() == 1) {
            uint64_t rem = 0; q.v.assign(a.v.size(), 0);
            for (size_t i = a.v.size(); i-- > 0;) {
                uint64_t cur = (rem << 32) | a.v[i];
                q.v[i] = (uint32_t)(cur / b.v[0]); rem = cur % b.v[0];
            }
            q.trim(); r = BigU(rem); return;
        }
        int s = __builtin_clz(b.v.back());
        BigU un = a.shl(s), vn = b.shl(s);
        const size_t n = vn.v.size(), m = a.v.size() - n;
        un.v.resize(a.v.size() + 1, 0);
        q.v.assign(m + 1, 0);
        for (size_t jj = m + 1; jj-- > 0;) {
            size_t j = jj;
            uint64_t num = ((uint64_t)un.v[j + n] << 32) | un.v[j + n - 1];
            uint64_t qhat = num / vn.v[n - 1], rhat = num % vn.v[n - 1];
            while (qhat >= (1ull << 32) || qhat * vn.v[n - 2] > ((rhat << 32) | un.v[j + n - 2])) {
                --qhat; rhat += vn.v[n - 1];
                if (rhat >= (1ull << 32)) break;
            }
            int64_t borrow = 0; uint64_t carry = 0;
            for (size_t i = 0; i < n; ++i) {
                uint64_t p = qhat * vn.v[i] + carry;
                carry = p >> 32;
                int64_t t = (int64_t)un.v[i + j] - borrow - (int64_t)(p & 0xffffffffu);
                un.v[i + j] = (uint32_t)t; borrow = t < 0;
            }
            int64_t t = (int64_t)un.v[j + n] - borrow - (int64_t)carry;
            un.v[j + n] = (uint32_t)t;
            if (t < 0) {
                --qhat; uint64_t c = 0;
                for (size_t i = 0; i < n; ++i) {
                    c += (uint64_t)un.v[i + j] + vn.v[i];
                    un.v[i + j] = (uint32_t)c; c >>= 32;
                }
                un.v[j + n] += (uint32_t)c;
            }
            q.v[j] = (uint32_t)qhat;
        }
        q.trim();
        un.v.resize(n); un.trim();
        r = un.shr(s);
    }
    BigU operator/(const BigU& o) const { BigU q, r; divmod(*this, o, q, r); return q; }
    BigU operator%(const BigU& o) const { BigU q, r; divmod(*this, o, q, r); return r; }

    static BigU gcd(BigU a, BigU b) {
        while (!b.is_zero()) { BigU t = a % b; a = b; b = t; }
        return a;
    }
    // a^-1 mod m; returns false if gcd(a, m) != 1
    static bool modinv(const BigU& a, const BigU& m, BigU& out) {
        // extended Euclid with coefficients tracked mod m as (value, sign)
        BigU r0 = m, r1 = a % m;
        BigU t0, t1(1); bool s0 = false, s1 = false;  // sign flags: true = negative
        while (!r1.is_zero()) {
            BigU q, r2; divmod(r0, r1, q, r2);
            // t2 = t0 - q*t1
            BigU qt = q * t1; BigU t2; bool s2;
            if (s0 == s1) {
                if (t0 >= qt) { t2 = t0 - qt; s2 = s0; } else { t2 = qt - t0; s2 = !s0; }
            } else { t2 = t0 + qt; s2 = s0; }
            r0 = r1; r1 = r2; t0 = t1; s0 = s1; t1 = t2; s1 = s2;
        }
        if (r0 != BigU(1)) return false;
        BigU t = t0 % m;
        out = (s0 && !t.is_zero()) ? m - t : t;
        return true;
    }
    static BigU modexp(const BigU& b, const BigU& e, const BigU& m) {
        BigU r(1); r = r % m; BigU base = b % m;
        for (size_t i = e.bitlen(); i-- > 0;) {
            r = (r * r) % m;
            if (e.bit(i)) r = (r * base) % m;
        }
        return r;
    }
    static BigU isqrt(const BigU& a) {
        if (a.is_zero()) return a;
        BigU x = pow2((a.bitlen() + 1) / 2);
        for (;;) {
            BigU y = (x + a / x).shr(1);
            if (y >= x) return x;
            x = y;
        }
    }
    static BigU factorial(unsigned n) { BigU r(1); for (unsigned i = 2; i <= n; ++i) r = r * BigU(i); return r; }

    std::string hex() const {
        if (v.empty()) return "0";
        static const char* d = "0123456789abcdef"; std::string s;
        for (size_t i = v.size(); i-- > 0;) for (int k = 28; k >= 0; k -= 4) s.push_back(d[(v[i] >> k) & 15]);
        size_t p = s.find_first_not_of('0'); return s.substr(p);
    }
};

// -n^-1 mod 2^32 for odd n
inline uint32_t mont_np0(uint32_t n0) {
    uint32_t x = n0;  // Newton: x = n0^-1 mod 2^32
    for (int i = 0; i < 5; ++i) x *= 2 - n0 * x;
    return (uint32_t)(0u - x);
}

}  // namespace pgpu
