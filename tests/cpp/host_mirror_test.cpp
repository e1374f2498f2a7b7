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
        } else {
            std::cerr << "unknown record " << kind << "\n";
            return 2;
        }
    }
    if (fails) { std::cerr << fails << " mismatches\n"; return 1; }
    std::cout << "host mirror ok\n";
    return 0;
}
