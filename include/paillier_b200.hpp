// paillier_b200.hpp -- header-only C++17 host mirror of the reference's interface for the batch path, above the
// C ABI of include/pgpu.h.  The reference is a Go package (compiled code, no Go toolchain in this image): this is
// the host layer a C++ caller links, with the reference's names, argument meaning and error behaviour:
//
//   paillier::PublicKey::EncryptWithRBatch        <- PublicKey.EncryptWithR            paillier.go:185-187,206-218
//   paillier::SecretKey::DecryptBatch             <- SecretKey.Decrypt                 paillier.go:292-303
//   paillier::SecretKey::EncryptWithRBatch        <- EncryptWithR on a SecretKey (CRT)  paillier.go:29-34,206-218
//   paillier::PublicKey::ConstMultBatch / AddBatch / SubPairs  <- operations.go:11-64
//   paillier::ThresholdSecretKey::PartialDecryptBatch / PartialDecryptionWithZKPBatch  <- thresholdkey.go:192-255
//   paillier::ThresholdPublicKey::VerifyProofBatch / CombinePartialDecryptionsBatch    <- thresholdkey.go:149-172,278-311
//
// Big integers cross this layer as paillier::Int = big-endian magnitude bytes, exactly what gmp.Int.Bytes() returns
// (zero is the empty string).  The layer only marshals to the fixed-width little-endian records of the C ABI; all
// arithmetic on batch items happens in libpaillier_b200.so on the GPU.  Errors: paillier::Error carries the PGPU_ERR_*
// code; the reference's error strings ("Threshold not meet", ...) are the message.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

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
}  // namespace detail

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
    // N x PublicKey.ConstMult with unsigned scalars (operations.go:58-64); k = 0 yields 1 like gmp's Exp
    std::vector<Ciphertext> ConstMultBatch(const std::vector<Ciphertext>& cts, const std::vector<Int>& ks) {
        if (cts.size() != ks.size()) throw Error(PGPU_ERR_ARG, "one scalar per ciphertext");
        size_t kb = 4;
        for (const Int& k : ks) kb = std::max(kb, (k.size() + 3) / 4 * 4);
        auto c = detail::to_records(values(cts), w_n2), k = detail::to_records(ks, kb);
        std::vector<uint8_t> o(cts.size() * w_n2);
        check(pgpu_const_mult(ctx_, cts.size(), c.data(), k.data(), kb, o.data()));
        auto out = wrap(detail::from_records(o, w_n2), EncLevelOne, RegularEncryption);
        for (size_t i = 0; i < out.size(); ++i) { out[i].Level = cts[i].Level; out[i].EncMethod = cts[i].EncMethod; }
        return out;
    }
    // PublicKey.Add(cts...) (operations.go:11-29)
    Ciphertext AddBatch(const std::vector<Ciphertext>& cts) {
        auto c = detail::to_records(values(cts), w_n2);
        std::vector<uint8_t> o(w_n2);
        check(pgpu_add_reduce(ctx_, cts.size(), cts.empty() ? nullptr : c.data(), o.data()));
        return wrap(detail::from_records(o, w_n2), EncLevelOne, MixedEncryption)[0];
    }
    // N x PublicKey.Sub(a_i, b_i) (operations.go:32-55)
    std::vector<Ciphertext> SubPairs(const std::vector<Ciphertext>& a, const std::vector<Ciphertext>& b) {
        if (a.size() != b.size()) throw Error(PGPU_ERR_ARG, "pairs");
        auto ra = detail::to_records(values(a), w_n2), rb = detail::to_records(values(b), w_n2);
        std::vector<uint8_t> o(a.size() * w_n2);
        check(pgpu_sub_pairs(ctx_, a.size(), ra.data(), rb.data(), o.data()));
        return wrap(detail::from_records(o, w_n2), EncLevelOne, MixedEncryption);
    }

    size_t w_n = 0, w_n2 = 0, w_n3 = 0;

protected:
    pgpu_ctx* ctx_ = nullptr;
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
    // N x SecretKey.Decrypt (paillier.go:292-303), level 1
    std::vector<Int> DecryptBatch(const std::vector<Ciphertext>& cts) {
        for (const auto& c : cts)
            if (c.Level != EncLevelOne) throw Error(PGPU_ERR_ARG, "DecryptBatch handles level-1 ciphertexts");
        auto c = detail::to_records(values(cts), w_n2);
        std::vector<uint8_t> m(cts.size() * w_n);
        check(pgpu_decrypt(ctx_, cts.size(), c.data(), m.data()));
        return detail::from_records(m, w_n);
    }
};

// ThresholdPublicKey (thresholdkey.go:26-32)
class ThresholdPublicKey : public PublicKey {
public:
    int TotalNumberOfDecryptionServers, Threshold;
    ThresholdPublicKey(const Int& n, int l, int w, const Int& v, const std::vector<Int>& vi, int device = 0, int id = 0, const Int* share = nullptr)
        : PublicKey(n, device), TotalNumberOfDecryptionServers(l), Threshold(w) {
        auto vk = detail::to_records(vi, w_n2);
        check(pgpu_ctx_set_threshold(ctx_, l, w, id, share ? share->data() : nullptr, share ? share->size() : 0, v.data(), v.size(),
                                     vi.empty() ? nullptr : vk.data()));
        check(pgpu_ctx_z_width(ctx_, &w_z));
    }
    // N x PartialDecryptionZKP.VerifyProof (thresholdkey.go:278-291); one server per batch
    std::vector<bool> VerifyProofBatch(const std::vector<PartialDecryptionZKP>& proofs) {
        if (proofs.empty()) return {};
        std::vector<Int> c, d, e, z;
        for (const auto& p : proofs) {
            if (p.ID != proofs[0].ID) throw Error(PGPU_ERR_ARG, "VerifyProofBatch: one server id per batch");
            c.push_back(p.C); d.push_back(p.Decryption); e.push_back(p.E); z.push_back(p.Z);
        }
        auto rc = detail::to_records(c, w_n2), rd = detail::to_records(d, w_n2), re = detail::to_records(e, 32), rz = detail::to_records(z, w_z);
        std::vector<uint8_t> ok(proofs.size());
        check(pgpu_pdec_zkp_verify(ctx_, proofs.size(), proofs[0].ID, rc.data(), rd.data(), re.data(), rz.data(), ok.data()));
        return std::vector<bool>(ok.begin(), ok.end());
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
        : ThresholdPublicKey(n, l, w, v, vi, device, id, &share), ID(id) {}
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
};

}  // namespace paillier
