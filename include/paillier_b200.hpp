// paillier_b200.hpp -- header-only C++17 host mirror of the reference's interface for the batch path, above the
// C ABI of include/pgpu.h.  The reference is a Go package (compiled code, no Go toolchain in this image): this is
// the host layer a C++ caller links, with the reference's names, argument meaning and error behaviour:
//
//   paillier::PublicKey::EncryptWithRBatch        <- PublicKey.EncryptWithR            paillier.go:185-187,206-218
//   paillier::SecretKey::DecryptBatch             <- SecretKey.Decrypt                 paillier.go:292-303
//   paillier::SecretKey::EncryptWithRBatch        <- EncryptWithR on a SecretKey (CRT)  paillier.go:29-34,206-218
//   paillier::PublicKey::ConstMultBatch / AddBatch / SubPairs  <- operations.go:11-64
//   paillier::PublicKey::EncryptWithRAtLevelBatch / AltEncryptWithRAtLevelBatch     <- paillier.go:206-238
//   paillier::PublicKey::RandomizeWithRBatch / NestedRandomizeWithBatch / NestedAddBatch / NestedSubBatch  <- operations.go:67-140
//   paillier::SecretKey::NestedDecryptBatch / ExtractRandonnessBatch / ProveDDLEQBatch, PublicKey::VerifyDDLEQProofBatch
//                                                   <- paillier.go:344-372, operations.go:75-91, ddleq.go:27-153
//   paillier::SafePrimeScan / MillerRabinBatch      <- safe_prime.go:170-278
//   paillier::ThresholdSecretKey::PartialDecryptBatch / PartialDecryptionWithZKPBatch  <- thresholdkey.go:192-255
//   paillier::ThresholdPublicKey::VerifyProofBatch / CombinePartialDecryptionsBatch / CombinePartialDecryptionsZKPBatch /
//     VerifyDecryptionBatch                                                             <- thresholdkey.go:149-189,278-311
//   paillier::ThresholdSecretKey::GetPublicKey / VerifyPartialDecryption                <- thresholdkey.go:213-222,258-275
//   paillier::DeviceBuffer / PinnedBytes, PublicKey::EncryptWithRDev / ConstMultDev / AddPairsDev / AddReduceDev / Sync,
//     SecretKey::DecryptDev, ThresholdSecretKey::PartialDecryptDev  <- ciphertexts kept on the GPU between operations.go calls
//   paillier::ThresholdGroup  <- one share-holder per GPU: thresholdkey.go:149-172,192-311 with one NCCL all-gather
//   the callers that draw their own randomness (paillier::RandomSource, default the OS CSPRNG) or take an argument list:
//   PublicKey::EncryptBatch / EncryptAtLevelBatch / NestedEncryptBatch / AltEncryptAtLevelBatch / EncryptZero*Batch /
//     EncryptOne*Batch / RandomizeBatch / NestedRandomizeBatch / SubBatch  <- paillier.go:192-203,244-289, operations.go:32-55,67-118
//
// Big integers cross this layer as paillier::Int = big-endian magnitude bytes, exactly what gmp.Int.Bytes() returns
// (zero is the empty string).  The layer only marshals to the fixed-width little-endian records of the C ABI; all
// arithmetic on batch items happens in libpaillier_b200.so on the GPU.  Errors: paillier::Error carries the PGPU_ERR_*
// code; the reference's error strings ("Threshold not meet", ...) are the message.
#pragma once
#include <cerrno>
#include <cstdint>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include <sys/random.h>
#include <sys/types.h>

#include "pgpu.h"

namespace paillier {

using Int = std::vector<uint8_t>;   // big-endian magnitude, no leading zeros (gmp.Int.Bytes())

inline Int from_hex(std::string h) {
    if (h.rfind("0x", 0) == 0) h = h.substr(2);
    if (h.size() % 2) h = "0" + h;
    Int out;
    for (size_t i = 0; i < h.size(); i += 2) out.push_back((uint8_t)std::stoul(h.substr(i, 2), nullptr, 16));
    size_t z = 0;
    while (z < out.size() && out[z] == 0) ++z;
    out.erase(out.begin(), out.begin() + z);
    return out;
}

inline std::string to_hex(const Int& v) {
    static const char* d = "0123456789abcdef";
    std::string s = "0x";
    if (v.empty()) return s + "0";
    bool lead = true;
    for (uint8_t b : v) {
        if (lead && (b >> 4) == 0) { s += d[b & 15]; lead = false; continue; }
        lead = false;
        s += d[b >> 4]; s += d[b & 15];
    }
    return s;
}

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

enum EncryptionLevel { EncLevelOne = 0, EncLevelTwo = 1 };                               // paillier.go:17-23
enum EncryptionMethod { RegularEncryption = 0, AlternativeEncryption = 1, MixedEncryption = 2 };   // paillier.go:29-39

struct Ciphertext {                                                                                 // paillier.go:65-69
    Int C; int Level = EncLevelOne; int EncMethod = RegularEncryption;
    std::vector<uint8_t> Bytes() const;                  // Ciphertext.Bytes, paillier.go:392-401 (encoding/gob stream)
};
struct DDLEQProofInstance { Int X, Y, Alpha, E, F; };                                               // ddleq.go:13-19
struct DDLEQProof { std::vector<DDLEQProofInstance> Instances; };                                   // ddleq.go:21-24
struct PartialDecryption { int ID = 0; Int Decryption; };                                           // thresholdkey.go:45-48
struct PartialDecryptionZKP { int ID = 0; Int Decryption, E, Z, C; };                               // thresholdkey.go:52-58

namespace detail {
// values -> `width`-byte little-endian records
inline std::vector<uint8_t> to_records(const std::vector<Int>& vals, size_t width) {
    std::vector<uint8_t> buf(vals.size() * width, 0);
    for (size_t i = 0; i < vals.size(); ++i) {
        const Int& v = vals[i];
        if (v.size() > width) throw Error(PGPU_ERR_ARG, "value wider than its record");
        for (size_t j = 0; j < v.size(); ++j) buf[i * width + j] = v[v.size() - 1 - j];
    }
    return buf;
}
inline std::vector<Int> from_records(const std::vector<uint8_t>& buf, size_t width) {
    std::vector<Int> out(width ? buf.size() / width : 0);
    for (size_t i = 0; i < out.size(); ++i) {
        size_t n = width;
        while (n > 0 && buf[i * width + n - 1] == 0) --n;
        out[i].resize(n);
        for (size_t j = 0; j < n; ++j) out[i][j] = buf[i * width + n - 1 - j];
    }
    return out;
}
// x mod 2^bits on a big-endian magnitude
inline void reduce_pow2(Int& x, unsigned bits) {
    const size_t keep = (bits + 7) / 8;
    if (x.size() > keep) x.erase(x.begin(), x.end() - keep);
    if (x.size() == keep && bits % 8) x[0] &= (uint8_t)((1u << (bits % 8)) - 1);
    size_t z = 0;
    while (z < x.size() && x[z] == 0) ++z;
    x.erase(x.begin(), x.begin() + z);
}
// big-endian magnitudes without leading zeros: a < b
inline bool less(const Int& a, const Int& b) { return a.size() != b.size() ? a.size() < b.size() : a < b; }
// a * b on big-endian magnitudes (schoolbook; key set-up only: n^2 for the range of a proof's r)
inline Int mul(const Int& a, const Int& b) {
    if (a.empty() || b.empty()) return {};
    std::vector<uint32_t> acc(a.size() + b.size(), 0);           // little-endian byte columns
    for (size_t i = 0; i < a.size(); ++i)
        for (size_t j = 0; j < b.size(); ++j) {
            const size_t k = i + j;
            acc[k] += (uint32_t)a[a.size() - 1 - i] * b[b.size() - 1 - j];
            if (acc[k] >> 24) { acc[k + 1] += acc[k] >> 8; acc[k] &= 0xff; }
        }
    Int out(acc.size());
    uint32_t carry = 0;
    for (size_t k = 0; k < acc.size(); ++k) { carry += acc[k]; out[acc.size() - 1 - k] = (uint8_t)carry; carry >>= 8; }
    size_t z = 0;
    while (z < out.size() && out[z] == 0) ++z;
    out.erase(out.begin(), out.begin() + z);
    return out;
}
}  // namespace detail

// Source of random bytes for the methods that draw their own randomness (the reference reads crypto/rand.Reader,
// utils.go:26-49): fills `len` bytes.  The default reads the operating system's CSPRNG.
using RandomSource = std::function<void(uint8_t*, size_t)>;
inline void os_random(uint8_t* out, size_t len) {
    while (len) {                                   // getrandom(2): the kernel's CSPRNG, no descriptor, no user-space buffer
        const ssize_t got = getrandom(out, len, 0);
        if (got < 0) {
            if (errno == EINTR) continue;
            throw Error(PGPU_ERR_STATE, "getrandom failed");
        }
        out += got; len -= (size_t)got;
    }
}
// uniform in [0, bound) by rejection, as crypto/rand.Int does (utils.go:26-33)
inline Int random_below(const Int& bound, const RandomSource& rnd) {
    if (bound.empty()) throw Error(PGPU_ERR_ARG, "random_below: bound must be positive");
    unsigned top_bits = 0;
    for (uint8_t t = bound[0]; t; t >>= 1) ++top_bits;
    for (;;) {
        Int r(bound.size());
        rnd(r.data(), r.size());
        r[0] &= (uint8_t)((1u << top_bits) - 1);
        size_t z = 0;
        while (z < r.size() && r[z] == 0) ++z;
        r.erase(r.begin(), r.begin() + z);
        if (detail::less(r, bound)) return r;
    }
}

// Wire format of a Ciphertext: Go's encoding/gob stream of struct{C *gmp.Int; Level, EncMethod int} written by a fresh
// encoder (paillier.go:374-401).  Restated from the encoding/gob specification and ncw/gmp's Int.GobEncode
// (version<<1|sign, big-endian magnitude); host-side marshalling only.  Same layout as paillier_b200/gobwire.py.
namespace gob {
inline void put_uint(std::vector<uint8_t>& o, uint64_t u) {
    if (u < 128) { o.push_back((uint8_t)u); return; }
    int n = 0;
    for (uint64_t t = u; t; t >>= 8) ++n;
    o.push_back((uint8_t)(256 - n));
    for (int i = n - 1; i >= 0; --i) o.push_back((uint8_t)(u >> (8 * i)));
}
inline void put_int(std::vector<uint8_t>& o, int64_t i) { put_uint(o, i < 0 ? ((uint64_t)~i << 1) | 1 : (uint64_t)i << 1); }
inline void put_string(std::vector<uint8_t>& o, const std::string& s) { put_uint(o, s.size()); o.insert(o.end(), s.begin(), s.end()); }
inline void put_named_id(std::vector<uint8_t>& o, const std::string& name, int64_t id) {   // {Name, Id} + end of struct
    o.push_back(1); put_string(o, name); o.push_back(1); put_int(o, id); o.push_back(0);
}
inline void put_message(std::vector<uint8_t>& o, const std::vector<uint8_t>& body) { put_uint(o, body.size()); o.insert(o.end(), body.begin(), body.end()); }

struct Reader {
    const uint8_t* d; size_t p, end;
    size_t left() const { return end - p; }
    const uint8_t* take(size_t n) { if (n > left()) throw Error(PGPU_ERR_ARG, "unexpected EOF"); const uint8_t* r = d + p; p += n; return r; }
    uint64_t uint() {
        uint8_t b = *take(1);
        if (b < 128) return b;
        size_t n = 256 - b;
        if (n > 8) throw Error(PGPU_ERR_ARG, "encoded unsigned integer out of range");
        const uint8_t* q = take(n);
        uint64_t v = 0;
        for (size_t i = 0; i < n; ++i) v = v << 8 | q[i];
        return v;
    }
    int64_t sint() { uint64_t u = uint(); return (u & 1) ? ~(int64_t)(u >> 1) : (int64_t)(u >> 1); }
    std::string string() { size_t n = (size_t)uint(); const uint8_t* q = take(n); return std::string(q, q + n); }
    std::pair<std::string, int64_t> named_id() {           // CommonType / fieldType: {Name string; Id typeId}
        std::string name; int64_t id = 0; int f = -1;
        for (;;) {
            uint64_t dl = uint();
            if (dl == 0) return {name, id};
            f += (int)dl;
            if (f == 0) name = string(); else if (f == 1) id = sint(); else throw Error(PGPU_ERR_ARG, "gob: unknown field in a type description");
        }
    }
};

inline std::vector<uint8_t> encode(const Ciphertext& ct, int64_t struct_id = 65) {
    const int64_t int_id = struct_id + 1;
    std::vector<uint8_t> out, b;
    put_int(b, -struct_id); b.push_back(3);                 // wireType.StructT
    b.push_back(1); put_named_id(b, "Ciphertext", struct_id);
    b.push_back(1); put_uint(b, 3);
    put_named_id(b, "C", int_id); put_named_id(b, "Level", 2); put_named_id(b, "EncMethod", 2);
    b.push_back(0); b.push_back(0);
    put_message(out, b); b.clear();
    put_int(b, -int_id); b.push_back(5);                    // wireType.GobEncoderT
    b.push_back(1); put_named_id(b, "Int", int_id);
    b.push_back(0); b.push_back(0);
    put_message(out, b); b.clear();
    put_int(b, struct_id);
    b.push_back(1); put_uint(b, ct.C.size() + 1); b.push_back(2); b.insert(b.end(), ct.C.begin(), ct.C.end());
    int last = 0;
    const int vals[2] = {ct.Level, ct.EncMethod};
    for (int i = 0; i < 2; ++i)
        if (vals[i] != 0) { put_uint(b, (uint64_t)(i + 1 - last)); put_int(b, vals[i]); last = i + 1; }
    b.push_back(0);
    put_message(out, b);
    return out;
}

inline Ciphertext decode(const uint8_t* data, size_t len) {
    if (len == 0) throw Error(PGPU_ERR_ARG, "no data provided");                         // paillier.go:377-379
    struct Type { bool is_struct = false; std::vector<std::pair<std::string, int64_t>> fields; };
    std::vector<std::pair<int64_t, Type>> types;
    auto find = [&](int64_t id) -> const Type* { for (auto& t : types) if (t.first == id) return &t.second; return nullptr; };
    Reader top{data, 0, len};
    while (top.left() > 0) {
        size_t n = (size_t)top.uint();
        top.take(n);
        Reader r{data, top.p - n, top.p};
        int64_t tid = r.sint();
        if (tid < 0) {
            if (find(-tid)) throw Error(PGPU_ERR_ARG, "gob: duplicate type received");
            Type t;
            uint64_t kind = r.uint();
            if (kind == 3) {
                t.is_struct = true;
                int f = -1;
                for (;;) {
                    uint64_t dl = r.uint();
                    if (dl == 0) break;
                    f += (int)dl;
                    if (f == 0) r.named_id();
                    else if (f == 1) { for (uint64_t k = r.uint(); k > 0; --k) t.fields.push_back(r.named_id()); }
                    else throw Error(PGPU_ERR_ARG, "gob: unknown field in structType");
                }
            } else if (kind >= 5 && kind <= 7) {
                uint64_t f = r.uint();
                if (f == 1) { r.named_id(); if (r.uint() != 0) throw Error(PGPU_ERR_ARG, "gob: unknown field in gobEncoderType"); }
                else if (f != 0) throw Error(PGPU_ERR_ARG, "gob: unknown field in gobEncoderType");
            } else throw Error(PGPU_ERR_ARG, "gob: type definition not used by a Ciphertext stream");
            if (r.uint() != 0) throw Error(PGPU_ERR_ARG, "gob: wireType with more than one kind");
            types.emplace_back(-tid, t);
            continue;
        }
        const Type* t = find(tid);
        if (!t || !t->is_struct) throw Error(PGPU_ERR_ARG, "gob: type mismatch: no fields matched compiling decoder for Ciphertext");
        Ciphertext ct;
        int f = -1;
        for (;;) {
            uint64_t dl = r.uint();
            if (dl == 0) break;
            f += (int)dl;
            if ((size_t)f >= t->fields.size()) throw Error(PGPU_ERR_ARG, "gob: field number out of range");
            const std::string& name = t->fields[f].first;
            const int64_t ft = t->fields[f].second;
            const Type* ftype = ft >= 64 ? find(ft) : nullptr;
            if ((ftype && !ftype->is_struct) || ft == 5 || ft == 6) {
                size_t bl = (size_t)r.uint();
                const uint8_t* q = r.take(bl);
                if (name == "C") {
                    if (!ftype) throw Error(PGPU_ERR_ARG, "gob: wrong type for field C");
                    if (bl > 0) {
                        if ((q[0] >> 1) != 1) throw Error(PGPU_ERR_ARG, "Int.GobDecode: encoding version not supported");
                        if (q[0] & 1) throw Error(PGPU_ERR_ARG, "negative ciphertext value");
                        size_t z = 1;
                        while (z < bl && q[z] == 0) ++z;
                        ct.C.assign(q + z, q + bl);
                    }
                }
            } else if (ft == 2) {
                int64_t v = r.sint();
                if (name == "Level") ct.Level = (int)v; else if (name == "EncMethod") ct.EncMethod = (int)v;
            } else if (ft == 1 || ft == 3 || ft == 4) {
                r.uint();
                if (name == "Level" || name == "EncMethod") throw Error(PGPU_ERR_ARG, "gob: wrong type for field " + name);
            } else throw Error(PGPU_ERR_ARG, "gob: cannot skip a field of this type");
        }
        if (r.left() != 0) throw Error(PGPU_ERR_ARG, "gob: extra data in message");
        return ct;
    }
    throw Error(PGPU_ERR_ARG, "unexpected EOF");
}
}  // namespace gob

inline std::vector<uint8_t> Ciphertext::Bytes() const { return gob::encode(*this); }
// PublicKey.NewCiphertextFromBytes (paillier.go:374-390); like the reference it needs no key material and does not range-check C
inline Ciphertext NewCiphertextFromBytes(const std::vector<uint8_t>& data) { return gob::decode(data.data(), data.size()); }

// Device memory on a key's device (pgpu_buf_*): what the *Dev methods below take, so that ciphertexts stay on the GPU between
// Encrypt / ConstMult / Add / Decrypt calls (the callers of operations.go:11-64).  Made by PublicKey::NewDeviceBuffer; free it
// (destructor) before the key it came from.
class DeviceBuffer {
public:
    DeviceBuffer(pgpu_ctx* ctx, size_t bytes) {
        const int rc = pgpu_buf_alloc(ctx, bytes, &buf_);
        if (rc != PGPU_OK) throw Error(rc, pgpu_last_error(ctx));
    }
    ~DeviceBuffer() { if (buf_) pgpu_buf_free(buf_); }
    DeviceBuffer(DeviceBuffer&& o) noexcept : buf_(o.buf_) { o.buf_ = nullptr; }
    DeviceBuffer& operator=(DeviceBuffer&& o) noexcept { if (this != &o) { if (buf_) pgpu_buf_free(buf_); buf_ = o.buf_; o.buf_ = nullptr; } return *this; }
    DeviceBuffer(const DeviceBuffer&) = delete;
    DeviceBuffer& operator=(const DeviceBuffer&) = delete;
    void* ptr() const { return pgpu_buf_ptr(buf_); }
    size_t size() const { return pgpu_buf_size(buf_); }
    // blocking copies of whole records at byte offset `off`
    void Upload(size_t off, const std::vector<uint8_t>& records) {
        const int rc = pgpu_buf_upload(buf_, off, records.data(), records.size());
        if (rc != PGPU_OK) throw Error(rc, pgpu_last_error(nullptr));
    }
    std::vector<uint8_t> Download(size_t off, size_t bytes) const {
        std::vector<uint8_t> out(bytes);
        const int rc = pgpu_buf_download(buf_, off, out.data(), bytes);
        if (rc != PGPU_OK) throw Error(rc, pgpu_last_error(nullptr));
        return out;
    }
private:
    pgpu_buf* buf_ = nullptr;
};

// Page-locked host memory (pgpu_host_alloc): the host-buffer entry points copy from / to it at the full PCIe rate and overlap
// the copies of one chunk with the kernels of the next (pageable memory works too, several times slower).
class PinnedBytes {
public:
    explicit PinnedBytes(size_t bytes) : size_(bytes) {
        const int rc = pgpu_host_alloc(bytes, &p_);
        if (rc != PGPU_OK) throw Error(rc, pgpu_last_error(nullptr));
    }
    ~PinnedBytes() { if (p_) pgpu_host_free(p_); }
    PinnedBytes(const PinnedBytes&) = delete;
    PinnedBytes& operator=(const PinnedBytes&) = delete;
    uint8_t* data() const { return static_cast<uint8_t*>(p_); }
    size_t size() const { return size_; }
private:
    void* p_ = nullptr;
    size_t size_ = 0;
};

// PublicKey{N} with g = n+1 (paillier.go:46-56,147); owns one engine context on `device`.
class PublicKey {
public:
    Int N;
    explicit PublicKey(const Int& n, int device = 0) : N(n) {
        check(pgpu_ctx_create(&ctx_, device, n.data(), n.size()));
        check(pgpu_ctx_widths(ctx_, &w_n, &w_n2, &w_n3));
    }
    virtual ~PublicKey() { if (ctx_) pgpu_ctx_destroy(ctx_); }
    PublicKey(const PublicKey&) = delete;
    PublicKey& operator=(const PublicKey&) = delete;

    Ciphertext NewCiphertextFromBytes(const std::vector<uint8_t>& data) const { return paillier::NewCiphertextFromBytes(data); }

    // N x PublicKey.EncryptWithR (paillier.go:185-187)
    std::vector<Ciphertext> EncryptWithRBatch(const std::vector<Int>& ms, const std::vector<Int>& rs) {
        if (ms.size() != rs.size()) throw Error(PGPU_ERR_ARG, "one r per plaintext");
        auto m = detail::to_records(ms, w_n), r = detail::to_records(rs, w_n);
        std::vector<uint8_t> c(ms.size() * w_n2);
        check(pgpu_encrypt_with_r(ctx_, ms.size(), m.data(), r.data(), c.data()));
        return wrap(detail::from_records(c, w_n2), EncLevelOne, RegularEncryption);
    }
    // operations.go:11-64 take n^(s+1) from the level of the (first) ciphertext (getModuliForLevel, paillier.go:403-414)
    int batch_level(const std::vector<Ciphertext>& cts, const char* what) const {
        const int level = cts.empty() ? (int)EncLevelOne : cts[0].Level;
        for (const Ciphertext& c : cts)
            if (c.Level != level) throw Error(PGPU_ERR_ARG, std::string(what) + ": one encryption level per batch");
        return level;
    }
    static int level_modsel(int level) { return level == EncLevelOne ? PGPU_MOD_N2 : PGPU_MOD_N3; }

    // N x PublicKey.ConstMult with unsigned scalars (operations.go:58-64) at the ciphertexts' level; k = 0 yields 1 like gmp's Exp
    std::vector<Ciphertext> ConstMultBatch(const std::vector<Ciphertext>& cts, const std::vector<Int>& ks) {
        if (cts.size() != ks.size()) throw Error(PGPU_ERR_ARG, "one scalar per ciphertext");
        const int level = batch_level(cts, "ConstMultBatch");
        const size_t w = cipher_width(level);
        size_t kb = 4;
        for (const Int& k : ks) kb = std::max(kb, (k.size() + 3) / 4 * 4);
        auto c = detail::to_records(values(cts), w), k = detail::to_records(ks, kb);
        std::vector<uint8_t> o(cts.size() * w);
        check(pgpu_modexp(ctx_, level_modsel(level), cts.size(), c.data(), k.data(), kb, o.data()));
        auto out = wrap(detail::from_records(o, w), (EncryptionLevel)level, RegularEncryption);
        for (size_t i = 0; i < out.size(); ++i) out[i].EncMethod = cts[i].EncMethod;
        return out;
    }
    // PublicKey.Add(cts...) (operations.go:11-29): modulus and level of cts[0]
    Ciphertext AddBatch(const std::vector<Ciphertext>& cts) {
        const int level = cts.empty() ? (int)EncLevelOne : cts[0].Level;
        const size_t w = cipher_width(level);
        auto c = detail::to_records(values(cts), w);
        std::vector<uint8_t> o(w);
        check(pgpu_add_reduce_at_level(ctx_, level == EncLevelOne ? 1 : 2, cts.size(), cts.empty() ? nullptr : c.data(), o.data()));
        return wrap(detail::from_records(o, w), (EncryptionLevel)level, MixedEncryption)[0];
    }
    // N x PublicKey.Sub(a_i, b_i) (operations.go:32-55): modulus and level of a_i, one level per batch
    std::vector<Ciphertext> SubPairs(const std::vector<Ciphertext>& a, const std::vector<Ciphertext>& b) {
        if (a.size() != b.size()) throw Error(PGPU_ERR_ARG, "pairs");
        const int level = batch_level(a, "SubPairs");
        const size_t w = cipher_width(level);
        auto ra = detail::to_records(values(a), w), rb = detail::to_records(values(b), w);
        std::vector<uint8_t> o(a.size() * w);
        if (level == EncLevelOne) {
            check(pgpu_sub_pairs(ctx_, a.size(), ra.data(), rb.data(), o.data()));
        } else if (!a.empty()) {
            std::vector<uint8_t> inv(rb.size());
            check(pgpu_modinv(ctx_, PGPU_MOD_N3, a.size(), rb.data(), inv.data()));
            check(pgpu_modmul(ctx_, PGPU_MOD_N3, a.size(), ra.data(), inv.data(), o.data()));
        }
        return wrap(detail::from_records(o, w), (EncryptionLevel)level, MixedEncryption);
    }
    // PublicKey.Sub(cts...) (operations.go:32-55): cts[0] * prod_{i>0} cts[i]^-1 modulo n^(s+1) of cts[0].Level.  The inverses
    // of the reference's loop are taken once, of the product of cts[1:] (the same canonical residue); a single argument comes
    // back as it is (:34).  Error(PGPU_ERR_NOT_INVERTIBLE) where ModInverse has no result.
    Ciphertext SubBatch(const std::vector<Ciphertext>& cts) {
        if (cts.empty()) throw Error(PGPU_ERR_ARG, "Sub needs at least one ciphertext");
        const int level = cts[0].Level;
        if (cts.size() == 1) return Ciphertext{cts[0].C, level, MixedEncryption};
        const size_t w = cipher_width(level);
        auto first = detail::to_records({cts[0].C}, w);
        auto rest = detail::to_records(values(std::vector<Ciphertext>(cts.begin() + 1, cts.end())), w);
        std::vector<uint8_t> prod(w), inv(w), o(w);
        check(pgpu_add_reduce_at_level(ctx_, level == EncLevelOne ? 1 : 2, cts.size() - 1, rest.data(), prod.data()));
        check(pgpu_modinv(ctx_, level_modsel(level), 1, prod.data(), inv.data()));
        check(pgpu_modmul(ctx_, level_modsel(level), 1, first.data(), inv.data(), o.data()));
        return wrap(detail::from_records(o, w), level, MixedEncryption)[0];
    }

    // N x PublicKey.Add(a_i, b_i) (operations.go:11-29): modulus and level of a_i, one level per batch
    std::vector<Ciphertext> AddPairs(const std::vector<Ciphertext>& a, const std::vector<Ciphertext>& b) {
        if (a.size() != b.size()) throw Error(PGPU_ERR_ARG, "pairs");
        const int level = batch_level(a, "AddPairs");
        const size_t w = cipher_width(level);
        auto ra = detail::to_records(values(a), w), rb = detail::to_records(values(b), w);
        std::vector<uint8_t> o(a.size() * w);
        check(pgpu_modmul(ctx_, level_modsel(level), a.size(), ra.data(), rb.data(), o.data()));
        return wrap(detail::from_records(o, w), (EncryptionLevel)level, MixedEncryption);
    }
    // Add(ConstMult(c_i, k_i)...) in one call: the encrypted dot product with 64-bit scalars
    Ciphertext DotProduct(const std::vector<Ciphertext>& cts, const std::vector<uint64_t>& ks) {
        if (cts.size() != ks.size()) throw Error(PGPU_ERR_ARG, "one scalar per ciphertext");
        auto c = detail::to_records(values(cts), w_n2);
        std::vector<uint8_t> o(w_n2);
        check(pgpu_dot_u64(ctx_, cts.size(), cts.empty() ? nullptr : c.data(), ks.data(), o.data()));
        return wrap(detail::from_records(o, w_n2), EncLevelOne, MixedEncryption)[0];
    }
    // r^n mod n^2 ahead of time, then N x EncryptWithR in two multiplications each (pgpu_encrypt_with_rn)
    virtual std::vector<Int> PrecomputeRnBatch(const std::vector<Int>& rs) {
        std::vector<Int> zeros(rs.size());
        std::vector<Int> out;
        for (auto& c : EncryptWithRBatch(zeros, rs)) out.push_back(std::move(c.C));
        return out;
    }
    std::vector<Ciphertext> EncryptWithRnBatch(const std::vector<Int>& ms, const std::vector<Int>& rns) {
        if (ms.size() != rns.size()) throw Error(PGPU_ERR_ARG, "one r^n per plaintext");
        auto m = detail::to_records(ms, w_n), r = detail::to_records(rns, w_n2);
        std::vector<uint8_t> c(ms.size() * w_n2);
        check(pgpu_encrypt_with_rn(ctx_, ms.size(), m.data(), r.data(), c.data()));
        return wrap(detail::from_records(c, w_n2), EncLevelOne, RegularEncryption);
    }
    // N x PublicKey.EncryptWithRAtLevel (paillier.go:206-218)
    virtual std::vector<Ciphertext> EncryptWithRAtLevelBatch(const std::vector<Int>& ms, const std::vector<Int>& rs, int level) {
        if (ms.size() != rs.size()) throw Error(PGPU_ERR_ARG, "one r per plaintext");
        auto m = detail::to_records(ms, plain_width(level)), r = detail::to_records(rs, w_n);
        std::vector<uint8_t> c(ms.size() * cipher_width(level));
        check(pgpu_encrypt_with_r_at_level(ctx_, level + 1, ms.size(), m.data(), r.data(), c.data()));
        return wrap(detail::from_records(c, cipher_width(level)), level, RegularEncryption);
    }
    // PublicKey.H, K = 2^k_bits (paillier.go:46-56): enables AltEncryptWithRAtLevelBatch
    void SetAltGenerator(const Int& H, unsigned k_bits) {
        check(pgpu_ctx_set_alt_generator(ctx_, H.data(), H.size(), k_bits));
        k_bits_ = k_bits;
    }
    // N x PublicKey.AltEncryptWithRAtLevel (paillier.go:221-238); like the reference (:228) the caller's r is reduced mod K
    std::vector<Ciphertext> AltEncryptWithRAtLevelBatch(const std::vector<Int>& ms, std::vector<Int>& rs, int level) {
        if (ms.size() != rs.size()) throw Error(PGPU_ERR_ARG, "one r per plaintext");
        if (k_bits_ == 0) throw Error(PGPU_ERR_STATE, "AltEncrypt needs PublicKey.H and K");
        for (Int& r : rs) detail::reduce_pow2(r, k_bits_);
        auto m = detail::to_records(ms, plain_width(level)), r = detail::to_records(rs, w_n);
        std::vector<uint8_t> c(ms.size() * cipher_width(level));
        check(pgpu_alt_encrypt_with_r_at_level(ctx_, level + 1, ms.size(), m.data(), r.data(), c.data()));
        return wrap(detail::from_records(c, cipher_width(level)), level, AlternativeEncryption);
    }
    // N x PublicKey.Randomize (operations.go:67-69) with the r of the fresh Encrypt(0) supplied.  Add takes the modulus from
    // ct.Level while Encrypt(0) is always a level-1 ciphertext: a level-2 ct is multiplied by r^n mod n^2 modulo n^3 (bit-exact
    // with the reference; that product is not an encryption of the same plaintext -- NestedRandomize is the level-2 form).
    std::vector<Ciphertext> RandomizeWithRBatch(const std::vector<Ciphertext>& cts, const std::vector<Int>& rs) {
        if (cts.size() != rs.size()) throw Error(PGPU_ERR_ARG, "one r per ciphertext");
        const int level = batch_level(cts, "RandomizeWithRBatch");
        if (level == EncLevelTwo) {
            auto zero = detail::to_records(values(EncryptWithRBatch(std::vector<Int>(rs.size()), rs)), w_n3);
            auto c = detail::to_records(values(cts), w_n3);
            std::vector<uint8_t> o(cts.size() * w_n3);
            if (!cts.empty()) check(pgpu_modmul(ctx_, PGPU_MOD_N3, cts.size(), c.data(), zero.data(), o.data()));
            return wrap(detail::from_records(o, w_n3), EncLevelTwo, MixedEncryption);
        }
        auto c = detail::to_records(values(cts), w_n2), r = detail::to_records(rs, w_n);
        std::vector<uint8_t> o(cts.size() * w_n2);
        check(pgpu_randomize_with_r(ctx_, cts.size(), c.data(), r.data(), o.data()));
        return wrap(detail::from_records(o, w_n2), EncLevelOne, MixedEncryption);
    }

    // ---- the callers that draw their own randomness -------------------------------------------------------------------
    // count x GetRandomNumberInMultiplicativeGroup(n) (utils.go:36-49): uniform below n, redrawn on 0 or gcd(n, r) != 1.  The
    // unit test of the whole batch is one batched ModInverse mod n on the GPU; only after a failure are the items checked
    // one by one to find the ones to redraw.
    std::vector<Int> DrawUnits(size_t count, const RandomSource& rnd = os_random) {
        std::vector<Int> rs(count);
        for (Int& r : rs) do r = random_below(N, rnd); while (r.empty());
        size_t w = 0;
        check(pgpu_ctx_mod_width(ctx_, PGPU_MOD_N, &w));
        while (count) {
            auto rec = detail::to_records(rs, w);
            std::vector<uint8_t> inv(rec.size());
            const int rc = pgpu_modinv(ctx_, PGPU_MOD_N, count, rec.data(), inv.data());
            if (rc == PGPU_OK) break;
            if (rc != PGPU_ERR_NOT_INVERTIBLE) check(rc);
            for (size_t i = 0; i < count; ++i) {
                const int one = pgpu_modinv(ctx_, PGPU_MOD_N, 1, rec.data() + i * w, inv.data());
                if (one == PGPU_ERR_NOT_INVERTIBLE) do rs[i] = random_below(N, rnd); while (rs[i].empty());
                else check(one);
            }
        }
        return rs;
    }
    // N x PublicKey.EncryptAtLevel (paillier.go:258-269) / Encrypt (:192-194) / NestedEncrypt (:200-203)
    std::vector<Ciphertext> EncryptAtLevelBatch(const std::vector<Int>& ms, int level, const RandomSource& rnd = os_random) {
        return EncryptWithRAtLevelBatch(ms, DrawUnits(ms.size(), rnd), level);
    }
    std::vector<Ciphertext> EncryptBatch(const std::vector<Int>& ms, const RandomSource& rnd = os_random) { return EncryptAtLevelBatch(ms, EncLevelOne, rnd); }
    std::vector<Ciphertext> NestedEncryptBatch(const std::vector<Int>& ms, const RandomSource& rnd = os_random) {
        return EncryptAtLevelBatch(values(EncryptAtLevelBatch(ms, EncLevelOne, rnd)), EncLevelTwo, rnd);
    }
    // N x PublicKey.AltEncryptAtLevel (paillier.go:244-255): r from Z*_n as there, reduced mod K by AltEncryptWithRAtLevel
    std::vector<Ciphertext> AltEncryptAtLevelBatch(const std::vector<Int>& ms, int level, const RandomSource& rnd = os_random) {
        auto rs = DrawUnits(ms.size(), rnd);
        return AltEncryptWithRAtLevelBatch(ms, rs, level);
    }
    // count x EncryptZero / EncryptOne (AtLevel) (paillier.go:272-289)
    std::vector<Ciphertext> EncryptZeroAtLevelBatch(size_t count, int level, const RandomSource& rnd = os_random) {
        return EncryptAtLevelBatch(std::vector<Int>(count), level, rnd);
    }
    std::vector<Ciphertext> EncryptOneAtLevelBatch(size_t count, int level, const RandomSource& rnd = os_random) {
        return EncryptAtLevelBatch(std::vector<Int>(count, Int{1}), level, rnd);
    }
    std::vector<Ciphertext> EncryptZeroBatch(size_t count, const RandomSource& rnd = os_random) { return EncryptZeroAtLevelBatch(count, EncLevelOne, rnd); }
    std::vector<Ciphertext> EncryptOneBatch(size_t count, const RandomSource& rnd = os_random) { return EncryptOneAtLevelBatch(count, EncLevelOne, rnd); }
    // N x PublicKey.Randomize (operations.go:67-69) with the r of each fresh Encrypt(0) drawn here
    std::vector<Ciphertext> RandomizeBatch(const std::vector<Ciphertext>& cts, const RandomSource& rnd = os_random) {
        return RandomizeWithRBatch(cts, DrawUnits(cts.size(), rnd));
    }
    // N x PublicKey.NestedRandomize (operations.go:96-118): a, b drawn from Z*_n (:105-106) and returned like there
    std::vector<Ciphertext> NestedRandomizeBatch(const std::vector<Ciphertext>& cts, std::vector<Int>& as, std::vector<Int>& bs,
                                                 const RandomSource& rnd = os_random) {
        as = DrawUnits(cts.size(), rnd);
        bs = DrawUnits(cts.size(), rnd);
        return NestedRandomizeWithBatch(cts, as, bs);
    }
    // N x PublicKey.NestedRandomize (operations.go:96-118) with (a, b) supplied
    std::vector<Ciphertext> NestedRandomizeWithBatch(const std::vector<Ciphertext>& cts, const std::vector<Int>& as, const std::vector<Int>& bs) {
        if (cts.size() != as.size() || cts.size() != bs.size()) throw Error(PGPU_ERR_ARG, "one (a, b) per ciphertext");
        for (const auto& c : cts)
            if (c.Level != EncLevelTwo) throw Error(PGPU_ERR_ARG, "can only homomorphically randomize doubly encrypted values");
        auto c = detail::to_records(values(cts), w_n3), a = detail::to_records(as, w_n), b = detail::to_records(bs, w_n);
        std::vector<uint8_t> o(cts.size() * w_n3);
        check(pgpu_nested_randomize_with(ctx_, cts.size(), c.data(), a.data(), b.data(), o.data()));
        return wrap(detail::from_records(o, w_n3), EncLevelTwo, RegularEncryption);
    }
    // N x PublicKey.NestedAdd / NestedSub (operations.go:121-140)
    std::vector<Ciphertext> NestedAddBatch(const std::vector<Ciphertext>& ct1, const std::vector<Ciphertext>& ct2) { return nested(pgpu_nested_add, ct1, ct2); }
    std::vector<Ciphertext> NestedSubBatch(const std::vector<Ciphertext>& ct1, const std::vector<Ciphertext>& ct2) { return nested(pgpu_nested_sub, ct1, ct2); }
    // N x PublicKey.VerifyDDLEQProof (ddleq.go:44-53); one instance count per batch
    std::vector<bool> VerifyDDLEQProofBatch(const std::vector<Ciphertext>& ct1, const std::vector<Ciphertext>& ct2, const std::vector<DDLEQProof>& proofs) {
        if (proofs.empty()) return {};
        if (ct1.size() != proofs.size() || ct2.size() != proofs.size()) throw Error(PGPU_ERR_ARG, "one statement per proof");
        const size_t secpar = proofs[0].Instances.size();
        if (secpar == 0) return std::vector<bool>(proofs.size(), true);
        std::vector<Int> x, y, al, e, f;
        for (const auto& p : proofs) {
            if (p.Instances.size() != secpar) throw Error(PGPU_ERR_ARG, "VerifyDDLEQProofBatch: one secpar per batch");
            for (const auto& i : p.Instances) { x.push_back(i.X); y.push_back(i.Y); al.push_back(i.Alpha); e.push_back(i.E); f.push_back(i.F); }
        }
        auto c1 = detail::to_records(values(ct1), w_n3), c2 = detail::to_records(values(ct2), w_n3);
        auto rx = detail::to_records(x, w_n), ry = detail::to_records(y, w_n);
        auto ra = detail::to_records(al, w_n3), re = detail::to_records(e, w_n2), rf = detail::to_records(f, w_n3);
        std::vector<uint8_t> ok(x.size());
        check(pgpu_ddleq_verify(ctx_, proofs.size(), (unsigned)secpar, c1.data(), c2.data(), rx.data(), ry.data(), ra.data(), re.data(), rf.data(), ok.data()));
        std::vector<bool> out(proofs.size(), true);
        for (size_t i = 0; i < ok.size(); ++i) if (!ok[i]) out[i / secpar] = false;
        return out;
    }

    // ---- device-resident batches: work is enqueued on the context's stream and NOT synchronised; Sync() waits for it ------
    DeviceBuffer NewDeviceBuffer(size_t bytes) { return DeviceBuffer(ctx_, bytes); }
    void Sync() { check(pgpu_ctx_sync(ctx_)); }
    // c[i] = EncryptWithR(m[i], r[i]) (paillier.go:185-187): n-width m and r, n2-width c
    void EncryptWithRDev(size_t count, const DeviceBuffer& m, const DeviceBuffer& r, DeviceBuffer& c) {
        need(m, count * w_n); need(r, count * w_n); need(c, count * w_n2);
        check(pgpu_encrypt_with_r_dev(ctx_, count, m.ptr(), r.ptr(), c.ptr()));
    }
    // out[i] = ConstMult(c[i], k[i]) (operations.go:58-64): k = k_bytes-wide little-endian unsigned scalars (multiple of 4)
    void ConstMultDev(size_t count, const DeviceBuffer& c, const DeviceBuffer& k, size_t k_bytes, DeviceBuffer& out) {
        need(c, count * w_n2); need(k, count * k_bytes); need(out, count * w_n2);
        check(pgpu_const_mult_dev(ctx_, count, c.ptr(), k.ptr(), k_bytes, out.ptr()));
    }
    // out[i] = Add(a[i], b[i]) (operations.go:11-29); out may be a or b
    void AddPairsDev(size_t count, const DeviceBuffer& a, const DeviceBuffer& b, DeviceBuffer& out) {
        need(a, count * w_n2); need(b, count * w_n2); need(out, count * w_n2);
        check(pgpu_add_pairs_dev(ctx_, count, a.ptr(), b.ptr(), out.ptr()));
    }
    // out = Add(c[0], ..., c[count-1]) as one tree reduction
    void AddReduceDev(size_t count, const DeviceBuffer& c, DeviceBuffer& out) {
        need(c, count * w_n2); need(out, w_n2);
        check(pgpu_add_reduce_dev(ctx_, count, c.ptr(), out.ptr()));
    }
    // the C ABI handle of this key's context (pgpu.h), for entry points this header does not wrap
    pgpu_ctx* Handle() const { return ctx_; }

    size_t w_n = 0, w_n2 = 0, w_n3 = 0;

protected:
    static void need(const DeviceBuffer& b, size_t bytes) {
        if (b.size() < bytes) throw Error(PGPU_ERR_ARG, "device buffer smaller than the batch");
    }
    pgpu_ctx* ctx_ = nullptr;
    unsigned k_bits_ = 0;
    size_t plain_width(int level) const {
        if (level != EncLevelOne && level != EncLevelTwo) throw Error(PGPU_ERR_ARG, "unsupported encryption level");
        return level == EncLevelOne ? w_n : w_n2;
    }
    size_t cipher_width(int level) const { return plain_width(level) == w_n ? w_n2 : w_n3; }
    template <class Fn>
    std::vector<Ciphertext> nested(Fn fn, const std::vector<Ciphertext>& ct1, const std::vector<Ciphertext>& ct2) {
        if (ct1.size() != ct2.size()) throw Error(PGPU_ERR_ARG, "pairs");
        for (size_t i = 0; i < ct1.size(); ++i)
            if (ct1[i].Level != EncLevelTwo || ct2[i].Level != EncLevelOne)
                throw Error(PGPU_ERR_ARG, "can only homomorphically add an encrypted value to a doubly encrypted value");
        auto a = detail::to_records(values(ct1), w_n3), b = detail::to_records(values(ct2), w_n2);
        std::vector<uint8_t> o(ct1.size() * w_n3);
        check(fn(ctx_, ct1.size(), a.data(), b.data(), o.data()));
        auto out = wrap(detail::from_records(o, w_n3), EncLevelTwo, RegularEncryption);
        for (size_t i = 0; i < out.size(); ++i) out[i].EncMethod = ct1[i].EncMethod;
        return out;
    }
    void check(int rc) const {
        if (rc != PGPU_OK) throw Error(rc, pgpu_last_error(ctx_));
    }
    static std::vector<Int> values(const std::vector<Ciphertext>& cts) {
        std::vector<Int> v;
        for (const auto& c : cts) v.push_back(c.C);
        return v;
    }
    static std::vector<Ciphertext> wrap(std::vector<Int> vals, int level, int method) {
        std::vector<Ciphertext> out;
        for (auto& v : vals) out.push_back(Ciphertext{std::move(v), level, method});
        return out;
    }
};

// SecretKey{PublicKey, Lambda} (paillier.go:59-62): the reference keeps Lambda = (p-1)(q-1) only.
class SecretKey : public PublicKey {
public:
    SecretKey(const Int& n, const Int& lambda, int device = 0) : PublicKey(n, device) {
        check(pgpu_ctx_set_secret_lambda(ctx_, lambda.data(), lambda.size()));
    }
    SecretKey(const Int& n, const Int& p, const Int& q, int device) : PublicKey(n, device) {
        check(pgpu_ctx_set_secret_pq(ctx_, p.data(), p.size(), q.data(), q.size()));
    }
    // N x EncryptWithR through the embedded PublicKey (paillier.go:29-34): the same ciphertexts as
    // PublicKey::EncryptWithRBatch, computed over p^2 and q^2 (pgpu_encrypt_with_r_sk)
    std::vector<Ciphertext> EncryptWithRBatch(const std::vector<Int>& ms, const std::vector<Int>& rs) {
        if (ms.size() != rs.size()) throw Error(PGPU_ERR_ARG, "one r per plaintext");
        auto m = detail::to_records(ms, w_n), r = detail::to_records(rs, w_n);
        std::vector<uint8_t> c(ms.size() * w_n2);
        check(pgpu_encrypt_with_r_sk(ctx_, ms.size(), m.data(), r.data(), c.data()));
        return wrap(detail::from_records(c, w_n2), EncLevelOne, RegularEncryption);
    }
    std::vector<Ciphertext> EncryptWithRAtLevelBatch(const std::vector<Int>& ms, const std::vector<Int>& rs, int level) override {
        if (ms.size() != rs.size()) throw Error(PGPU_ERR_ARG, "one r per plaintext");
        auto m = detail::to_records(ms, plain_width(level)), r = detail::to_records(rs, w_n);
        std::vector<uint8_t> c(ms.size() * cipher_width(level));
        check(pgpu_encrypt_with_r_at_level_sk(ctx_, level + 1, ms.size(), m.data(), r.data(), c.data()));
        return wrap(detail::from_records(c, cipher_width(level)), level, RegularEncryption);
    }
    std::vector<Int> PrecomputeRnBatch(const std::vector<Int>& rs) override {
        std::vector<Int> zeros(rs.size()), out;
        for (auto& c : EncryptWithRBatch(zeros, rs)) out.push_back(std::move(c.C));
        return out;
    }
    // m[i] = Decrypt(c[i]) (paillier.go:292-303, CRT over p^2, q^2) on device buffers: n2-width c, n-width m; enqueued, see Sync()
    void DecryptDev(size_t count, const DeviceBuffer& c, DeviceBuffer& m) {
        need(c, count * w_n2); need(m, count * w_n);
        check(pgpu_decrypt_dev(ctx_, count, c.ptr(), m.ptr()));
    }
    // N x SecretKey.Decrypt (paillier.go:292-340); one encryption level per batch
    std::vector<Int> DecryptBatch(const std::vector<Ciphertext>& cts) {
        if (cts.empty()) return {};
        const int level = cts[0].Level;
        for (const auto& c : cts)
            if (c.Level != level) throw Error(PGPU_ERR_ARG, "DecryptBatch: one encryption level per batch");
        auto c = detail::to_records(values(cts), cipher_width(level));
        std::vector<uint8_t> m(cts.size() * plain_width(level));
        if (level == EncLevelOne) check(pgpu_decrypt(ctx_, cts.size(), c.data(), m.data()));
        else check(pgpu_decrypt_at_level(ctx_, level + 1, cts.size(), c.data(), m.data()));
        return detail::from_records(m, plain_width(level));
    }
    // N x SecretKey.DecryptNestedCiphertextLayer (paillier.go:360-372)
    std::vector<Ciphertext> DecryptNestedCiphertextLayerBatch(const std::vector<Ciphertext>& cts) {
        for (const auto& c : cts)
            if (c.Level == EncLevelOne) throw Error(PGPU_ERR_ARG, "no nested ciphertexts to recover");
        return wrap(DecryptBatch(cts), EncLevelOne, MixedEncryption);
    }
    // N x SecretKey.NestedDecrypt (paillier.go:344-356): an inner value 0 decrypts to 0 (:350-354)
    std::vector<Int> NestedDecryptBatch(const std::vector<Ciphertext>& cts) {
        auto inner = DecryptNestedCiphertextLayerBatch(cts);
        std::vector<Ciphertext> nz;
        std::vector<size_t> at;
        for (size_t i = 0; i < inner.size(); ++i)
            if (!inner[i].C.empty()) { nz.push_back(inner[i]); at.push_back(i); }
        auto vals = DecryptBatch(nz);
        std::vector<Int> out(cts.size());
        for (size_t k = 0; k < at.size(); ++k) out[at[k]] = std::move(vals[k]);
        return out;
    }
    // N x SecretKey.ExtractRandonness (operations.go:75-91); one level per batch
    std::vector<Int> ExtractRandonnessBatch(const std::vector<Ciphertext>& cts) {
        if (cts.empty()) return {};
        const int level = cts[0].Level;
        auto c = detail::to_records(values(cts), cipher_width(level));
        std::vector<uint8_t> o(cts.size() * w_n);
        check(pgpu_extract_randomness(ctx_, level + 1, cts.size(), c.data(), o.data()));
        return detail::from_records(o, w_n);
    }
    // N x SecretKey.ProveDDLEQ (ddleq.go:27-40): xs[i][j], ys[i][j] in Z*_n = the randomness of instance j of statement i
    // (:71-79).  Throws PGPU_ERR_ARG where the reference panics on wrong inputs (:67-69).
    std::vector<DDLEQProof> ProveDDLEQBatch(unsigned secpar, const std::vector<Ciphertext>& ct1, const std::vector<Ciphertext>& ct2,
                                            const std::vector<Int>& as, const std::vector<Int>& bs,
                                            const std::vector<std::vector<Int>>& xs, const std::vector<std::vector<Int>>& ys) {
        const size_t count = ct1.size();
        std::vector<DDLEQProof> out(count);
        if (secpar == 0 || count == 0) return out;
        if (ct2.size() != count || as.size() != count || bs.size() != count || xs.size() != count || ys.size() != count)
            throw Error(PGPU_ERR_ARG, "ProveDDLEQBatch: one (ct2, a, b, xs, ys) per statement");
        std::vector<Int> fx, fy;
        for (size_t i = 0; i < count; ++i) {
            if (xs[i].size() != secpar || ys[i].size() != secpar) throw Error(PGPU_ERR_ARG, "ProveDDLEQBatch: secpar values of x and y per statement");
            fx.insert(fx.end(), xs[i].begin(), xs[i].end()); fy.insert(fy.end(), ys[i].begin(), ys[i].end());
        }
        auto c1 = detail::to_records(values(ct1), w_n3), c2 = detail::to_records(values(ct2), w_n3);
        auto a = detail::to_records(as, w_n), b = detail::to_records(bs, w_n), x = detail::to_records(fx, w_n), y = detail::to_records(fy, w_n);
        const size_t total = count * secpar;
        std::vector<uint8_t> al(total * w_n3), e(total * w_n2), f(total * w_n3);
        check(pgpu_ddleq_prove(ctx_, count, secpar, c1.data(), c2.data(), a.data(), b.data(), x.data(), y.data(), al.data(), e.data(), f.data()));
        auto A = detail::from_records(al, w_n3), E = detail::from_records(e, w_n2), F = detail::from_records(f, w_n3);
        for (size_t k = 0; k < total; ++k)
            out[k / secpar].Instances.push_back(DDLEQProofInstance{fx[k], fy[k], std::move(A[k]), std::move(E[k]), std::move(F[k])});
        return out;
    }
};

// One iteration of the safe-prime search (safe_prime.go:170-278) for each `raw` byte string (ceil((p_bits-1)/8) bytes each,
// as read from the reference's io.Reader): candidate q (after the sieve's delta search), p = 2q+1 and whether the pair is accepted.
struct SafePrimeCandidate { Int p, q; bool ok = false; };
inline std::vector<SafePrimeCandidate> SafePrimeScan(unsigned p_bits, const std::vector<uint8_t>& raw, int device = 0) {
    const size_t rec = (p_bits - 1 + 7) / 8;
    if (rec == 0 || raw.size() % rec) throw Error(PGPU_ERR_ARG, "raw: whole records of ceil((p_bits-1)/8) bytes");
    const size_t count = raw.size() / rec, w = 4 * (p_bits <= 1024 ? 32 : p_bits <= 1536 ? 48 : 64);
    std::vector<uint8_t> p(count * w), q(count * w), ok(count);
    int rc = pgpu_safe_prime_scan(device, p_bits, count, raw.data(), p.data(), q.data(), ok.data(), nullptr);
    if (rc != PGPU_OK) throw Error(rc, pgpu_primes_last_error());
    auto P = detail::from_records(p, w), Q = detail::from_records(q, w);
    std::vector<SafePrimeCandidate> out(count);
    for (size_t i = 0; i < count; ++i) out[i] = SafePrimeCandidate{std::move(P[i]), std::move(Q[i]), ok[i] != 0};
    return out;
}
// big.Int.ProbablyPrime stand-in: `rounds` Miller-Rabin tests (bases 2, 3, 5, ...) on odd candidates of exactly `bits` bits
inline std::vector<bool> MillerRabinBatch(unsigned bits, const std::vector<Int>& cand, unsigned rounds = 20, int device = 0) {
    const size_t w = 4 * (bits <= 1024 ? 32 : bits <= 1536 ? 48 : 64);
    auto c = detail::to_records(cand, w);
    std::vector<uint8_t> ok(cand.size());
    int rc = pgpu_miller_rabin(device, bits, cand.size(), c.data(), rounds, ok.data(), nullptr);
    if (rc != PGPU_OK) throw Error(rc, pgpu_primes_last_error());
    return std::vector<bool>(ok.begin(), ok.end());
}

// ThresholdPublicKey (thresholdkey.go:26-32)
class ThresholdPublicKey : public PublicKey {
public:
    int TotalNumberOfDecryptionServers, Threshold;
    Int VerificationKey;
    std::vector<Int> VerificationKeys;
    ThresholdPublicKey(const Int& n, int l, int w, const Int& v, const std::vector<Int>& vi, int device = 0, int id = 0, const Int* share = nullptr)
        : PublicKey(n, device), TotalNumberOfDecryptionServers(l), Threshold(w), VerificationKey(v), VerificationKeys(vi) {
        auto vk = detail::to_records(vi, w_n2);
        check(pgpu_ctx_set_threshold(ctx_, l, w, id, share ? share->data() : nullptr, share ? share->size() : 0, v.data(), v.size(),
                                     vi.empty() ? nullptr : vk.data()));
        check(pgpu_ctx_z_width(ctx_, &w_z));
    }
    // N x PartialDecryptionZKP.VerifyProof (thresholdkey.go:278-291); one server per batch.  A value wider than its record
    // (E is a SHA-256 digest, Z < 2^(8 w_z) for every r < n^2, c and c_i residues mod n^2) cannot come from an honest
    // prover: such a proof is answered false, as the reference's VerifyProof would, instead of failing the batch.
    std::vector<bool> VerifyProofBatch(const std::vector<PartialDecryptionZKP>& proofs) {
        if (proofs.empty()) return {};
        std::vector<Int> c, d, e, z;
        std::vector<bool> fits;
        for (const auto& p : proofs) {
            if (p.ID != proofs[0].ID) throw Error(PGPU_ERR_ARG, "VerifyProofBatch: one server id per batch");
            const bool f = p.C.size() <= w_n2 && p.Decryption.size() <= w_n2 && p.E.size() <= 32 && p.Z.size() <= w_z;
            fits.push_back(f);
            c.push_back(f ? p.C : Int{}); d.push_back(f ? p.Decryption : Int{}); e.push_back(f ? p.E : Int{}); z.push_back(f ? p.Z : Int{});
        }
        auto rc = detail::to_records(c, w_n2), rd = detail::to_records(d, w_n2), re = detail::to_records(e, 32), rz = detail::to_records(z, w_z);
        std::vector<uint8_t> ok(proofs.size());
        check(pgpu_pdec_zkp_verify(ctx_, proofs.size(), proofs[0].ID, rc.data(), rd.data(), re.data(), rz.data(), ok.data()));
        std::vector<bool> out(proofs.size());
        for (size_t i = 0; i < out.size(); ++i) out[i] = ok[i] != 0 && fits[i];
        return out;
    }
    // N x CombinePartialDecryptionsZKP (thresholdkey.go:164-172): shares[j] = server j's batch, same ciphertext order.  As in
    // the reference the proofs filter PER CIPHERTEXT: ciphertext i is combined from the servers whose proof for i verifies.
    // Where fewer than Threshold remain the reference answers "Threshold not meet" for that ciphertext: with item_ok == nullptr
    // the call throws Error(PGPU_ERR_THRESHOLD) if that happens to any of them; otherwise (*item_ok)[i] tells which plaintexts
    // are valid (the others are empty).
    std::vector<Int> CombinePartialDecryptionsZKPBatch(const std::vector<std::vector<PartialDecryptionZKP>>& shares,
                                                       std::vector<bool>* item_ok = nullptr) {
        if (shares.empty()) throw Error(PGPU_ERR_THRESHOLD, "Threshold not meet");
        const size_t count = shares[0].size();
        std::vector<int> ids;
        std::vector<Int> flat;
        std::vector<uint8_t> ok;
        for (const auto& s : shares) {
            if (s.size() != count) throw Error(PGPU_ERR_ARG, "CombinePartialDecryptionsZKPBatch: one proof per ciphertext and server");
            ids.push_back(count ? s[0].ID : 0);
            for (bool v : VerifyProofBatch(s)) ok.push_back(v ? 1 : 0);
            for (const auto& p : s) flat.push_back(p.Decryption.size() <= w_n2 ? p.Decryption : Int{});
        }
        auto d = detail::to_records(flat, w_n2);
        std::vector<uint8_t> m(count * w_n, 0), flags(count, 0);
        const int rc = pgpu_combine_verified(ctx_, count, (int)ids.size(), ids.data(), d.empty() ? nullptr : d.data(), ok.empty() ? nullptr : ok.data(),
                                             m.empty() ? nullptr : m.data(), flags.empty() ? nullptr : flags.data());
        if (rc != PGPU_OK && !(rc == PGPU_ERR_THRESHOLD && item_ok)) check(rc);
        if (item_ok) item_ok->assign(flags.begin(), flags.end());
        return detail::from_records(m, w_n);
    }
    // N x ThresholdPublicKey.VerifyDecryption (thresholdkey.go:175-189): throws Error(PGPU_ERR_ARG) with the reference's
    // error strings, or the combine's error when too few valid shares remain
    void VerifyDecryptionBatch(const std::vector<Int>& encryptedMessages, const std::vector<Int>& decryptedMessages,
                               const std::vector<std::vector<PartialDecryptionZKP>>& shares) {
        for (const auto& s : shares) {
            if (s.size() != encryptedMessages.size()) throw Error(PGPU_ERR_ARG, "The encrypted message is not the same than the one in the shares");
            for (size_t i = 0; i < s.size(); ++i)
                if (s[i].C != encryptedMessages[i]) throw Error(PGPU_ERR_ARG, "The encrypted message is not the same than the one in the shares");
        }
        if (CombinePartialDecryptionsZKPBatch(shares) != decryptedMessages)
            throw Error(PGPU_ERR_ARG, "The decrypted message is not the same than the one in the shares");
    }
    // N x CombinePartialDecryptions (thresholdkey.go:149-161): shares[j] = server j's batch, same ciphertext order.
    // Throws Error(PGPU_ERR_THRESHOLD, "Threshold not meet" / duplicate server) like the reference's errors (:77-89).
    std::vector<Int> CombinePartialDecryptionsBatch(const std::vector<std::vector<PartialDecryption>>& shares) {
        std::vector<int> ids;
        std::vector<Int> flat;
        const size_t count = shares.empty() ? 0 : shares[0].size();
        for (const auto& s : shares) {
            ids.push_back(s.empty() ? 0 : s[0].ID);
            for (const auto& pd : s) flat.push_back(pd.Decryption);
        }
        auto d = detail::to_records(flat, w_n2);
        std::vector<uint8_t> m(count * w_n);
        check(pgpu_combine(ctx_, count, (int)ids.size(), ids.data(), d.empty() ? nullptr : d.data(), m.empty() ? nullptr : m.data()));
        return detail::from_records(m, w_n);
    }
    size_t w_z = 0;
};

// ThresholdSecretKey (thresholdkey.go:38-42)
class ThresholdSecretKey : public ThresholdPublicKey {
public:
    int ID;
    ThresholdSecretKey(const Int& n, int l, int w, const Int& v, const std::vector<Int>& vi, int id, const Int& share, int device = 0)
        : ThresholdPublicKey(n, l, w, v, vi, device, id, &share), ID(id), device_(device) {}
    // ThresholdSecretKey.PublicKey (thresholdkey.go:213-222): the key without ID and Share, on a context of its own
    std::unique_ptr<ThresholdPublicKey> GetPublicKey() const {
        return std::make_unique<ThresholdPublicKey>(N, TotalNumberOfDecryptionServers, Threshold, VerificationKey, VerificationKeys, device_);
    }
    // ThresholdSecretKey.VerifyPartialDecryption (thresholdkey.go:258-275): encrypt a random m < n, prove its partial
    // decryption (r < n^2, :233), verify the proof; `count` such checks run as one batch.  Throws Error(PGPU_ERR_ARG, "Invalid share").
    void VerifyPartialDecryption(size_t count = 1, const RandomSource& rnd = os_random) {
        const Int n2 = detail::mul(N, N);
        std::vector<Int> ms, rs, cs;
        for (size_t i = 0; i < count; ++i) { ms.push_back(random_below(N, rnd)); rs.push_back(random_below(n2, rnd)); }
        for (auto& c : EncryptBatch(ms, rnd)) cs.push_back(std::move(c.C));
        for (bool ok : VerifyProofBatch(PartialDecryptionWithZKPBatch(cs, rs)))
            if (!ok) throw Error(PGPU_ERR_ARG, "Invalid share");
    }
    // out[i] = PartialDecrypt(c[i]) (thresholdkey.go:192-201) on device buffers of n2-width records; enqueued, see Sync()
    void PartialDecryptDev(size_t count, const DeviceBuffer& c, DeviceBuffer& out) {
        need(c, count * w_n2); need(out, count * w_n2);
        check(pgpu_partial_decrypt_dev(ctx_, count, c.ptr(), out.ptr()));
    }
    // N x ThresholdSecretKey.PartialDecrypt (thresholdkey.go:192-201)
    std::vector<PartialDecryption> PartialDecryptBatch(const std::vector<Int>& cs) {
        auto c = detail::to_records(cs, w_n2);
        std::vector<uint8_t> o(cs.size() * w_n2);
        check(pgpu_partial_decrypt(ctx_, cs.size(), c.data(), o.data()));
        std::vector<PartialDecryption> out;
        for (auto& v : detail::from_records(o, w_n2)) out.push_back(PartialDecryption{ID, std::move(v)});
        return out;
    }
    // N x PartialDecryptionWithZKP (thresholdkey.go:225-255); rs = the r in [0, n^2) the reference draws at :233
    std::vector<PartialDecryptionZKP> PartialDecryptionWithZKPBatch(const std::vector<Int>& cs, const std::vector<Int>& rs) {
        auto c = detail::to_records(cs, w_n2), r = detail::to_records(rs, w_n2);
        std::vector<uint8_t> d(cs.size() * w_n2), e(cs.size() * 32), z(cs.size() * w_z);
        check(pgpu_pdec_zkp_prove(ctx_, cs.size(), c.data(), r.data(), d.data(), e.data(), z.data()));
        auto D = detail::from_records(d, w_n2), E = detail::from_records(e, 32), Z = detail::from_records(z, w_z);
        std::vector<PartialDecryptionZKP> out;
        for (size_t i = 0; i < cs.size(); ++i) out.push_back(PartialDecryptionZKP{ID, D[i], E[i], Z[i], cs[i]});
        return out;
    }

private:
    int device_ = 0;
};

// Threshold decryption with one share-holder per GPU of this process (pgpu_multi_*, BASELINE config 4): PartialDecrypt (+ proofs)
// of every ciphertext on every device, one NCCL all-gather, proof verification and Combine of a ciphertext slice per device
// (thresholdkey.go:149-172,192-311).  Takes one ThresholdSecretKey per device, all shares of the same key; the keys must
// outlive the group.
class ThresholdGroup {
public:
    explicit ThresholdGroup(const std::vector<ThresholdSecretKey*>& keys) : keys_(keys) {
        if (keys.empty()) throw Error(PGPU_ERR_ARG, "ThresholdGroup: at least one share-holder");
        std::vector<pgpu_ctx*> raw;
        for (auto* k : keys) raw.push_back(k->Handle());
        const int rc = pgpu_multi_create(&m_, raw.data(), (int)raw.size());
        if (rc != PGPU_OK) throw Error(rc, pgpu_multi_last_error(nullptr));
    }
    ~ThresholdGroup() { if (m_) pgpu_multi_destroy(m_); }
    ThresholdGroup(const ThresholdGroup&) = delete;
    ThresholdGroup& operator=(const ThresholdGroup&) = delete;
    int Size() const { return pgpu_multi_size(m_); }
    // Plaintexts of cs.  zkp_rs is empty (PartialDecrypt + CombinePartialDecryptions) or holds, per share-holder, one r in
    // [0, n^2) per ciphertext (PartialDecryptionWithZKP + CombinePartialDecryptionsZKP; the reference draws r at
    // thresholdkey.go:233).  Where too few proofs verify the reference answers "Threshold not meet" for that ciphertext: with
    // item_ok == nullptr the call throws Error(PGPU_ERR_THRESHOLD) if that happens to any; otherwise (*item_ok)[i] tells which
    // plaintexts are valid.
    std::vector<Int> Decrypt(const std::vector<Int>& cs, const std::vector<std::vector<Int>>& zkp_rs = {}, std::vector<bool>* item_ok = nullptr) {
        const size_t w_n = keys_[0]->w_n, w_n2 = keys_[0]->w_n2, count = cs.size();
        auto c = detail::to_records(cs, w_n2);
        std::vector<std::vector<uint8_t>> rrec;
        std::vector<const void*> rp;
        if (!zkp_rs.empty()) {
            if (zkp_rs.size() != keys_.size()) throw Error(PGPU_ERR_ARG, "one vector of randomness per share-holder");
            for (const auto& rs : zkp_rs) {
                if (rs.size() != count) throw Error(PGPU_ERR_ARG, "one r per ciphertext and share-holder");
                rrec.push_back(detail::to_records(rs, w_n2));
            }
            for (const auto& r : rrec) rp.push_back(r.data());
        }
        std::vector<uint8_t> m(count * w_n, 0), flags(count, 0);
        const int rc = pgpu_multi_threshold_round(m_, count, c.data(), rp.empty() ? nullptr : rp.data(), m.data(), flags.data());
        if (rc != PGPU_OK && !(rc == PGPU_ERR_THRESHOLD && item_ok)) throw Error(rc, pgpu_multi_last_error(m_));
        if (item_ok) item_ok->assign(flags.begin(), flags.end());
        return detail::from_records(m, w_n);
    }
    // device milliseconds of the last Decrypt per phase (max over the devices): PartialDecrypt, proofs, all-gather, VerifyProof, Combine
    std::vector<float> PhasesMs() const {
        std::vector<float> out(5, 0.f);
        pgpu_multi_last_phases_ms(m_, out.data());
        return out;
    }
private:
    std::vector<ThresholdSecretKey*> keys_;
    pgpu_multi* m_ = nullptr;
};

}  // namespace paillier
