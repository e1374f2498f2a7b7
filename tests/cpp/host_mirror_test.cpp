// Exercises include/paillier_b200.hpp (the C++ host mirror) against vectors given on stdin; driven by
// tests/test_cpp_host_mirror.py.  Input: whitespace-separated tokens, all integers in hex.
//   paillier  N lambda count  (m r c k ck) x count  add_all
//   threshold N l w V l x vi  w x (id share)  count  (c) x count  (m) x count
// Exit status 0 when every recomputed value matches.
#include <algorithm>
#include <iostream>
#include <string>
#include <vector>

#include "paillier_b200.hpp"

using namespace paillier;

static Int rd() { std::string s; std::cin >> s; return from_hex(s); }
static int fails = 0;
static void expect(bool ok, const char* what) { if (!ok) { std::cerr << "MISMATCH: " << what << "\n"; ++fails; } }

int main() {
    std::string kind;
    while (std::cin >> kind) {
        if (kind == "paillier") {
            Int N = rd(), lambda = rd();
            size_t count; std::cin >> count;
            std::vector<Int> ms, rs, cs, ks, cks;
            for (size_t i = 0; i < count; ++i) { ms.push_back(rd()); rs.push_back(rd()); cs.push_back(rd()); ks.push_back(rd()); cks.push_back(rd()); }
            Int add_all = rd();
            SecretKey sk(N, lambda);
            auto cts = sk.EncryptWithRBatch(ms, rs);
            for (size_t i = 0; i < count; ++i) expect(cts[i].C == cs[i], "SecretKey::EncryptWithRBatch");
            auto pub = static_cast<PublicKey&>(sk).EncryptWithRBatch(ms, rs);
            for (size_t i = 0; i < count; ++i) expect(pub[i].C == cs[i], "PublicKey::EncryptWithRBatch");
            expect(sk.DecryptBatch(cts) == ms, "DecryptBatch");
            auto cm = sk.ConstMultBatch(cts, ks);
            for (size_t i = 0; i < count; ++i) expect(cm[i].C == cks[i], "ConstMultBatch");
            expect(sk.AddBatch(cts).C == add_all, "AddBatch");
            auto back = sk.SubPairs({sk.AddBatch({cts[0], cts[1]})}, {cts[1]});
            expect(back[0].C == cts[0].C, "SubPairs(Add(a, b), b) == a");
            expect(sk.EncryptWithRBatch({}, {}).empty(), "empty batch");
        } else if (kind == "threshold") {
            Int N = rd();
            int l, w; std::cin >> l >> w;
            Int V = rd();
            std::vector<Int> vi;
            for (int i = 0; i < l; ++i) vi.push_back(rd());
            std::vector<std::pair<int, Int>> shares;
            for (int i = 0; i < w; ++i) { int id; std::cin >> id; shares.emplace_back(id, rd()); }
            size_t count; std::cin >> count;
            std::vector<Int> cs, ms, zr;
            for (size_t i = 0; i < count; ++i) cs.push_back(rd());
            for (size_t i = 0; i < count; ++i) ms.push_back(rd());
            for (size_t i = 0; i < count; ++i) zr.push_back(rd());
            std::vector<std::vector<PartialDecryption>> parts;
            ThresholdPublicKey tk(N, l, w, V, vi);
            for (auto& s : shares) {
                ThresholdSecretKey tsk(N, l, w, V, vi, s.first, s.second);
                parts.push_back(tsk.PartialDecryptBatch(cs));
                auto zk = tsk.PartialDecryptionWithZKPBatch(cs, zr);
                for (size_t i = 0; i < count; ++i) expect(zk[i].Decryption == parts.back()[i].Decryption, "ZKP decryption == PartialDecrypt");
                auto ok = tk.VerifyProofBatch(zk);
                expect(std::all_of(ok.begin(), ok.end(), [](bool b) { return b; }), "VerifyProofBatch accepts");
                zk[0].Z.back() ^= 1;
                expect(!tk.VerifyProofBatch(zk)[0], "VerifyProofBatch rejects a tampered Z");
            }
            expect(tk.CombinePartialDecryptionsBatch(parts) == ms, "CombinePartialDecryptionsBatch");
            parts.pop_back();
            try { tk.CombinePartialDecryptionsBatch(parts); expect(false, "Threshold not meet must throw"); }
            catch (const Error& e) { expect(e.code == PGPU_ERR_THRESHOLD && std::string(e.what()) == "Threshold not meet", "Threshold not meet"); }
        } else if (kind == "level2") {
            // level 2, alternative encryption, randomness extraction, nested ops and DDLEQ of one golden case
            Int N = rd(), lambda = rd(), H = rd();
            unsigned k_bits; std::cin >> k_bits;
            size_t count; std::cin >> count;
            std::vector<Int> m1, m2, r2, c2;
            for (size_t i = 0; i < count; ++i) { m1.push_back(rd()); m2.push_back(rd()); r2.push_back(rd()); c2.push_back(rd()); }
            size_t na; std::cin >> na;
            std::vector<Int> ar, ac1, ac2;
            for (size_t i = 0; i < na; ++i) { ar.push_back(rd()); ac1.push_back(rd()); ac2.push_back(rd()); }
            SecretKey sk(N, lambda);
            sk.SetAltGenerator(H, k_bits);
            auto ct2 = sk.EncryptWithRAtLevelBatch(m2, r2, EncLevelTwo);
            auto pub2 = static_cast<PublicKey&>(sk).PublicKey::EncryptWithRAtLevelBatch(m2, r2, EncLevelTwo);
            for (size_t i = 0; i < count; ++i) {
                expect(ct2[i].C == c2[i] && ct2[i].Level == EncLevelTwo, "SecretKey::EncryptWithRAtLevelBatch(level 2)");
                expect(pub2[i].C == c2[i], "PublicKey::EncryptWithRAtLevelBatch(level 2)");
            }
            expect(sk.DecryptBatch(ct2) == m2, "DecryptBatch(level 2)");
            std::vector<Int> r2head(r2.begin(), r2.begin() + 2);
            expect(sk.ExtractRandonnessBatch({ct2[0], ct2[1]}) == r2head, "ExtractRandonnessBatch(level 2)");
            std::vector<Int> ms3(m1.begin(), m1.begin() + na), m23(m2.begin(), m2.begin() + na), ar1 = ar, ar2 = ar;
            auto a1 = sk.AltEncryptWithRAtLevelBatch(ms3, ar1, EncLevelOne);
            auto a2 = sk.AltEncryptWithRAtLevelBatch(m23, ar2, EncLevelTwo);
            for (size_t i = 0; i < na; ++i) {
                expect(a1[i].C == ac1[i] && a1[i].EncMethod == AlternativeEncryption, "AltEncryptWithRAtLevelBatch(level 1)");
                expect(a2[i].C == ac2[i], "AltEncryptWithRAtLevelBatch(level 2)");
            }
            // offline/online EncryptWithR and Randomize: Randomize(E(m; r), s) = E(m; r*s)
            auto c1 = sk.EncryptWithRBatch(m1, r2);
            auto viaPool = sk.EncryptWithRnBatch(m1, sk.PrecomputeRnBatch(r2));
            for (size_t i = 0; i < count; ++i) expect(viaPool[i].C == c1[i].C, "EncryptWithRnBatch(PrecomputeRnBatch)");
            std::vector<Int> ones(count, Int{1});
            auto same = sk.RandomizeWithRBatch(c1, ones);
            for (size_t i = 0; i < count; ++i) expect(same[i].C == c1[i].C, "RandomizeWithRBatch(r = 1)");
            expect(sk.DecryptBatch(sk.RandomizeWithRBatch(c1, r2)) == m1, "Decrypt(Randomize(c)) == m");
            expect(sk.DecryptBatch(sk.AddPairs(c1, c1)) == sk.DecryptBatch(sk.ConstMultBatch(c1, std::vector<Int>(count, Int{2}))), "AddPairs(c, c) ~ ConstMult(c, 2)");
            expect(sk.DotProduct(c1, std::vector<uint64_t>(count, 1)).C == sk.AddBatch(c1).C, "DotProduct(k = 1) == AddBatch");
            // nested: [[m]] = Enc2(Enc1(m)); NestedDecrypt peels both layers; NestedAdd/NestedSub act on the inner plaintext
            std::vector<Int> innerC;
            for (auto& c : c1) innerC.push_back(c.C);
            auto outer = sk.EncryptWithRAtLevelBatch(innerC, r2, EncLevelTwo);
            expect(sk.NestedDecryptBatch(outer) == m1, "NestedDecryptBatch");
            auto back = sk.NestedSubBatch(sk.NestedAddBatch(outer, c1), c1);
            expect(sk.NestedDecryptBatch(back) == m1, "NestedSub(NestedAdd(x, c), c)");
            expect(sk.NestedDecryptBatch(sk.NestedRandomizeWithBatch(outer, r2, r2)) == m1, "NestedDecrypt(NestedRandomize(x)) == m");
            // DDLEQ golden proof
            Int d1 = rd(), d2 = rd(), da = rd(), db = rd();
            size_t secpar; std::cin >> secpar;
            std::vector<Int> dx, dy, dal, de, df;
            for (size_t i = 0; i < secpar; ++i) { dx.push_back(rd()); dy.push_back(rd()); dal.push_back(rd()); de.push_back(rd()); df.push_back(rd()); }
            Ciphertext s1{d1, EncLevelTwo, RegularEncryption}, s2{d2, EncLevelTwo, RegularEncryption};
            auto proofs = sk.ProveDDLEQBatch((unsigned)secpar, {s1}, {s2}, {da}, {db}, {dx}, {dy});
            for (size_t i = 0; i < secpar; ++i) {
                const auto& in = proofs[0].Instances[i];
                expect(in.Alpha == dal[i] && in.E == de[i] && in.F == df[i], "ProveDDLEQBatch");
            }
            expect(sk.VerifyDDLEQProofBatch({s1}, {s2}, proofs)[0], "VerifyDDLEQProofBatch accepts");
            proofs[0].Instances[0].F.back() ^= 1;
            expect(!sk.VerifyDDLEQProofBatch({s1}, {s2}, proofs)[0], "VerifyDDLEQProofBatch rejects a tampered F");
            // wire format round trip (paillier_test.go:140-156)
            auto rt = sk.NewCiphertextFromBytes(ct2[0].Bytes());
            expect(rt.C == ct2[0].C && rt.Level == EncLevelTwo && rt.EncMethod == RegularEncryption, "Bytes / NewCiphertextFromBytes");
        } else if (kind == "safeprime") {
            unsigned bits; size_t count; std::cin >> bits >> count;
            std::vector<uint8_t> raw;
            std::vector<Int> qs; std::vector<int> oks;
            for (size_t i = 0; i < count; ++i) {
                std::string h; std::cin >> h;
                for (size_t j = 0; j + 1 < h.size(); j += 2) raw.push_back((uint8_t)std::stoul(h.substr(j, 2), nullptr, 16));
                qs.push_back(rd()); int ok; std::cin >> ok; oks.push_back(ok);
            }
            auto got = SafePrimeScan(bits, raw);
            std::vector<Int> accepted;
            for (size_t i = 0; i < count; ++i) {
                expect(got[i].q == qs[i] && (int)got[i].ok == oks[i], "SafePrimeScan");
                if (got[i].ok) accepted.push_back(got[i].q);
            }
            if (!accepted.empty()) {
                auto mr = MillerRabinBatch(bits - 1, accepted);
                for (bool b : mr) expect(b, "MillerRabinBatch on accepted q");
            }
        } else {
            std::cerr << "unknown record " << kind << "\n";
            return 2;
        }
    }
    if (fails) { std::cerr << fails << " mismatches\n"; return 1; }
    std::cout << "host mirror ok\n";
    return 0;
}
