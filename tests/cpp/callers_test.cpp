// Exercises the C++ host mirror's callers that draw their own randomness or take a variadic argument list
// (include/paillier_b200.hpp: DrawUnits, EncryptBatch ... VerifyPartialDecryption, SubBatch,
// CombinePartialDecryptionsZKPBatch, VerifyDecryptionBatch); driven by tests/test_cpp_host_mirror.py (host helpers, no GPU)
// and tests/test_gpu_zz_callers.py (GPU).  Input: whitespace-separated tokens, integers in hex.
//   mul a b                 -> prints a*b
//   below bound             -> prints 64 draws of random_below(bound) from the deterministic source below
//   osrandom bound          -> the same from os_random
//   paillier N lambda H k_bits
//   threshold N l w V (vi) x l (id share) x l
//   device N lambda
//   group N V v1 share          (a 1-of-1 threshold key)
#include <algorithm>
#include <iostream>
#include <string>
#include <vector>

#include "paillier_b200.hpp"

using namespace paillier;

static Int rd() { std::string s; std::cin >> s; return from_hex(s); }
static int fails = 0;
static void expect(bool ok, const char* what) { if (!ok) { std::cerr << "MISMATCH: " << what << "\n"; ++fails; } }

// xorshift64*: a reproducible byte source for the tests (NOT for keys)
static uint64_t state = 0x9e3779b97f4a7c15ull;
static void test_random(uint8_t* out, size_t len) {
    for (size_t i = 0; i < len; ++i) {
        state ^= state >> 12; state ^= state << 25; state ^= state >> 27;
        out[i] = (uint8_t)((state * 0x2545f4914f6cdd1dull) >> 56);
    }
}

template <class T>
static std::vector<Int> values(const std::vector<T>& cts) { std::vector<Int> v; for (auto& c : cts) v.push_back(c.C); return v; }

int main() {
    std::string kind;
    while (std::cin >> kind) {
        if (kind == "mul") {
            Int a = rd(), b = rd();
            std::cout << to_hex(detail::mul(a, b)) << "\n";
        } else if (kind == "below") {
            Int bound = rd();
            for (int i = 0; i < 64; ++i) {
                Int r = random_below(bound, test_random);
                expect(detail::less(r, bound), "random_below < bound");
                std::cout << to_hex(r) << (i == 63 ? "\n" : " ");
            }
        } else if (kind == "osrandom") {                      // the default source: the operating system's CSPRNG
            Int bound = rd();
            for (int i = 0; i < 64; ++i) std::cout << to_hex(random_below(bound, os_random)) << (i == 63 ? "\n" : " ");
        } else if (kind == "paillier") {
            Int N = rd(), lambda = rd(), H = rd();
            unsigned k_bits; std::cin >> k_bits;
            SecretKey sk(N, lambda);
            sk.SetAltGenerator(H, k_bits);
            const Int n2 = detail::mul(N, N);
            std::vector<Int> ms, m2;
            for (int i = 0; i < 5; ++i) { ms.push_back(random_below(N, test_random)); m2.push_back(random_below(n2, test_random)); }
            for (int pass = 0; pass < 2; ++pass) {            // the deterministic source, then the operating system's
                const RandomSource rnd = pass ? RandomSource(os_random) : RandomSource(test_random);
                auto units = sk.DrawUnits(7, rnd);
                for (auto& u : units) expect(!u.empty() && detail::less(u, N), "DrawUnits in [1, n)");
                auto cts = sk.EncryptBatch(ms, rnd);
                expect(sk.DecryptBatch(cts) == ms, "EncryptBatch round trip");
                expect(sk.EncryptBatch(ms, rnd)[0].C != cts[0].C, "fresh randomness per call");
                expect(sk.DecryptBatch(sk.EncryptAtLevelBatch(m2, EncLevelTwo, rnd)) == m2, "EncryptAtLevelBatch level 2");
                auto nested = sk.NestedEncryptBatch(ms, rnd);
                expect(nested[0].Level == EncLevelTwo && sk.NestedDecryptBatch(nested) == ms, "NestedEncryptBatch round trip");
                auto alt = sk.AltEncryptAtLevelBatch(ms, EncLevelOne, rnd);
                expect(alt[0].EncMethod == AlternativeEncryption && sk.DecryptBatch(alt) == ms, "AltEncryptAtLevelBatch level 1");
                expect(sk.DecryptBatch(sk.AltEncryptAtLevelBatch(m2, EncLevelTwo, rnd)) == m2, "AltEncryptAtLevelBatch level 2");
                auto rz = sk.RandomizeBatch(cts, rnd);
                expect(values(rz) != values(cts) && sk.DecryptBatch(rz) == ms, "RandomizeBatch keeps the plaintext");
                std::vector<Int> as, bs;
                auto nr = sk.NestedRandomizeBatch(nested, as, bs, rnd);
                expect(as.size() == ms.size() && bs.size() == ms.size() && sk.NestedDecryptBatch(nr) == ms, "NestedRandomizeBatch keeps the plaintext");
                expect(values(sk.NestedRandomizeWithBatch(nested, as, bs)) == values(nr), "NestedRandomizeBatch returns its (a, b)");
                // at level 2 the reference's Randomize multiplies by a LEVEL-1 Encrypt(0) modulo n^3 (Add takes the modulus of
                // ct.Level): bit-exact with that product, whatever it decrypts to
                auto l2 = sk.EncryptAtLevelBatch(m2, EncLevelTwo, rnd);
                auto rs2 = sk.DrawUnits(l2.size(), rnd);
                auto l2r = sk.RandomizeWithRBatch(l2, rs2);
                auto want = sk.AddPairs(l2, static_cast<PublicKey&>(sk).EncryptWithRBatch(std::vector<Int>(rs2.size()), rs2));
                expect(l2r[0].Level == EncLevelTwo && values(l2r) != values(l2) && values(l2r) == values(want), "RandomizeWithRBatch at level 2");
                expect(sk.RandomizeBatch(l2, rnd).size() == l2.size(), "RandomizeBatch at level 2");
            }
            const Int zero, one{1};
            expect(sk.DecryptBatch(sk.EncryptZeroBatch(3)) == std::vector<Int>(3, zero), "EncryptZeroBatch");
            expect(sk.DecryptBatch(sk.EncryptOneBatch(3)) == std::vector<Int>(3, one), "EncryptOneBatch");
            expect(sk.DecryptBatch(sk.EncryptZeroAtLevelBatch(2, EncLevelTwo)) == std::vector<Int>(2, zero), "EncryptZeroAtLevelBatch");
            expect(sk.DecryptBatch(sk.EncryptOneAtLevelBatch(2, EncLevelTwo)) == std::vector<Int>(2, one), "EncryptOneAtLevelBatch");
            expect(sk.EncryptBatch({}).empty(), "empty batch");
            // Sub(cts...): Sub(Add(a, b, c), b, c) == a at both levels; one argument comes back as it is
            auto cts = sk.EncryptBatch(ms, test_random);
            auto sum = sk.AddBatch({cts[0], cts[1], cts[2]});
            expect(sk.SubBatch({sum, cts[1], cts[2]}).C == cts[0].C, "SubBatch(Add(a, b, c), b, c) == a");
            expect(sk.SubBatch({cts[3]}).C == cts[3].C && sk.SubBatch({cts[3]}).EncMethod == MixedEncryption, "SubBatch of one argument");
            auto l2 = sk.EncryptAtLevelBatch(m2, EncLevelTwo, test_random);
            auto sum2 = sk.AddBatch({l2[0], l2[1], l2[2]});
            auto back2 = sk.SubBatch({sum2, l2[1], l2[2]});
            expect(back2.Level == EncLevelTwo && back2.C == l2[0].C, "SubBatch at level 2");
            try { sk.SubBatch({cts[0], Ciphertext{N, EncLevelOne, RegularEncryption}}); expect(false, "Sub of a non-unit must throw"); }
            catch (const Error& e) { expect(e.code == PGPU_ERR_NOT_INVERTIBLE, "Sub of a non-unit: PGPU_ERR_NOT_INVERTIBLE"); }
        } else if (kind == "threshold") {
            Int N = rd();
            int l, w; std::cin >> l >> w;
            Int V = rd();
            std::vector<Int> vi;
            for (int i = 0; i < l; ++i) vi.push_back(rd());
            std::vector<std::unique_ptr<ThresholdSecretKey>> keys;
            for (int i = 0; i < l; ++i) { int id; std::cin >> id; Int share = rd(); keys.push_back(std::make_unique<ThresholdSecretKey>(N, l, w, V, vi, id, share)); }
            for (auto& k : keys) k->VerifyPartialDecryption(2, test_random);
            keys[0]->VerifyPartialDecryption();
            auto tk = keys[0]->GetPublicKey();
            expect(tk->Threshold == w && tk->VerificationKeys == vi, "GetPublicKey");
            const Int n2 = detail::mul(N, N);
            std::vector<Int> ms, cs;
            for (int i = 0; i < 4; ++i) ms.push_back(random_below(N, test_random));
            for (auto& c : tk->EncryptBatch(ms, test_random)) cs.push_back(c.C);
            std::vector<std::vector<PartialDecryptionZKP>> shares;
            for (auto& k : keys) {
                std::vector<Int> rs;
                for (size_t i = 0; i < cs.size(); ++i) rs.push_back(random_below(n2, test_random));
                shares.push_back(k->PartialDecryptionWithZKPBatch(cs, rs));
            }
            expect(tk->CombinePartialDecryptionsZKPBatch(shares) == ms, "CombinePartialDecryptionsZKPBatch");
            tk->VerifyDecryptionBatch(cs, ms, shares);
            // the proofs filter per ciphertext: spoil l - w + 1 proofs of ciphertext 1 (too few remain) and one proof of ciphertext 2
            auto spoiled = shares;
            for (int j = 0; j < l - w + 1; ++j) spoiled[j][1].E.back() ^= 1;
            spoiled[0][2].Z.back() ^= 1;
            std::vector<bool> item_ok;
            auto got = tk->CombinePartialDecryptionsZKPBatch(spoiled, &item_ok);
            expect(item_ok == std::vector<bool>({true, false, l - 1 >= w, true}), "per-ciphertext verdicts");
            expect(got[0] == ms[0] && got[3] == ms[3] && (l - 1 < w || got[2] == ms[2]), "ciphertexts with enough valid proofs still decrypt");
            try { tk->CombinePartialDecryptionsZKPBatch(spoiled); expect(false, "Threshold not meet must throw"); }
            catch (const Error& e) { expect(e.code == PGPU_ERR_THRESHOLD, "Threshold not meet"); }
            try { tk->VerifyDecryptionBatch(ms, ms, shares); expect(false, "VerifyDecryptionBatch must compare the ciphertexts"); }
            catch (const Error& e) { expect(std::string(e.what()) == "The encrypted message is not the same than the one in the shares", "VerifyDecryption error string"); }
            // an out-of-range Z is a false verdict, not an error
            auto wide = shares[0];
            wide[0].Z = Int(tk->w_z + 1, 0xff);
            auto verdicts = tk->VerifyProofBatch(wide);
            expect(!verdicts[0] && verdicts[1], "a proof wider than its record verifies false");
            // a corrupted share fails its own check
            Int share_bad = from_hex("0x1234567");
            ThresholdSecretKey bad(N, l, w, V, vi, 1, share_bad);
            try { bad.VerifyPartialDecryption(1, test_random); expect(false, "Invalid share must throw"); }
            catch (const Error& e) { expect(std::string(e.what()) == "Invalid share", "Invalid share"); }
        } else if (kind == "device") {
            // Encrypt -> ConstMult -> Add -> Decrypt with the ciphertexts staying on the GPU (DeviceBuffer), against the host-buffer batch calls
            Int N = rd(), lambda = rd();
            SecretKey sk(N, lambda);
            const size_t count = 6, kb = 8;
            std::vector<Int> ms, ks;
            for (size_t i = 0; i < count; ++i) { ms.push_back(random_below(N, test_random)); ks.push_back(random_below(from_hex("0xffffffffffffffff"), test_random)); }
            auto rs = sk.DrawUnits(count, test_random);
            auto cts = sk.EncryptWithRBatch(ms, rs);
            auto sum = sk.AddBatch(sk.ConstMultBatch(cts, ks));
            auto plain = sk.DecryptBatch({sum});
            auto m = sk.NewDeviceBuffer(count * sk.w_n), r = sk.NewDeviceBuffer(count * sk.w_n), k = sk.NewDeviceBuffer(count * kb);
            auto c = sk.NewDeviceBuffer(count * sk.w_n2), c2 = sk.NewDeviceBuffer(count * sk.w_n2), total = sk.NewDeviceBuffer(sk.w_n2), out = sk.NewDeviceBuffer(sk.w_n);
            expect(m.size() == count * sk.w_n && m.ptr() != nullptr, "DeviceBuffer size / ptr");
            m.Upload(0, detail::to_records(ms, sk.w_n)); r.Upload(0, detail::to_records(rs, sk.w_n)); k.Upload(0, detail::to_records(ks, kb));
            sk.EncryptWithRDev(count, m, r, c);
            sk.ConstMultDev(count, c, k, kb, c2);
            sk.AddReduceDev(count, c2, total);
            sk.DecryptDev(1, total, out);
            sk.Sync();
            std::vector<Int> cvals;
            for (auto& ct : cts) cvals.push_back(ct.C);
            expect(detail::from_records(c.Download(0, count * sk.w_n2), sk.w_n2) == cvals, "EncryptWithRDev == EncryptWithRBatch");
            expect(detail::from_records(total.Download(0, sk.w_n2), sk.w_n2)[0] == sum.C, "ConstMultDev + AddReduceDev == AddBatch(ConstMultBatch)");
            expect(detail::from_records(out.Download(0, sk.w_n), sk.w_n) == plain, "DecryptDev == DecryptBatch");
            sk.AddPairsDev(count, c, c2, c2);
            sk.Sync();
            auto pairs = sk.AddPairs(cts, sk.ConstMultBatch(cts, ks));
            std::vector<Int> pvals;
            for (auto& ct : pairs) pvals.push_back(ct.C);
            expect(detail::from_records(c2.Download(0, count * sk.w_n2), sk.w_n2) == pvals, "AddPairsDev in place == AddPairs");
            try { sk.EncryptWithRDev(count + 1, m, r, c); expect(false, "a batch larger than its buffers must throw"); }
            catch (const Error& e) { expect(e.code == PGPU_ERR_ARG, "device buffer smaller than the batch"); }
            DeviceBuffer moved = std::move(total);
            expect(moved.size() == sk.w_n2, "DeviceBuffer is movable");
            PinnedBytes pin(1 << 16);
            expect(pin.data() != nullptr && pin.size() == (1 << 16), "PinnedBytes");
            auto recs = detail::to_records(ms, sk.w_n);
            std::copy(recs.begin(), recs.end(), pin.data());
            m.Upload(0, std::vector<uint8_t>(pin.data(), pin.data() + recs.size()));
        } else if (kind == "group") {
            // one share-holder per device of this process (pgpu_multi_*): here a 1-of-1 key on device 0, a one-rank communicator
            Int N = rd(), V = rd(), v1 = rd(), share = rd();
            ThresholdSecretKey tsk(N, 1, 1, V, {v1}, 1, share);
            ThresholdGroup grp({&tsk});
            expect(grp.Size() == 1, "ThresholdGroup size");
            const Int n2 = detail::mul(N, N);
            std::vector<Int> ms, cs, rs;
            for (int i = 0; i < 5; ++i) { ms.push_back(random_below(N, test_random)); rs.push_back(random_below(n2, test_random)); }
            for (auto& c : tsk.EncryptBatch(ms, test_random)) cs.push_back(c.C);
            expect(grp.Decrypt(cs) == ms, "ThresholdGroup::Decrypt without proofs");
            std::vector<bool> item_ok;
            expect(grp.Decrypt(cs, {rs}, &item_ok) == ms && item_ok == std::vector<bool>(5, true), "ThresholdGroup::Decrypt with proofs");
            expect(grp.PhasesMs().size() == 5, "PhasesMs");
            try { grp.Decrypt(cs, {rs, rs}); expect(false, "one vector of randomness per share-holder"); }
            catch (const Error& e) { expect(e.code == PGPU_ERR_ARG, "one vector of randomness per share-holder"); }
        } else {
            std::cerr << "unknown section " << kind << "\n";
            return 2;
        }
    }
    if (fails) return 1;
    std::cout << "callers ok\n";
    return 0;
}
