// batch_gpu_ops.go -- second half of the cgo binding (see batch_gpu.go): level 2, alternative encryption,
// pairwise homomorphic operations, nested operations, proof verification and the generic Exp.  Marshalling
// only; NOT COMPILED in this repository's image (no Go toolchain) -- the same symbols are exercised through
// ctypes (paillier_b200/api.py) and through the C++ mirror (include/paillier_b200.hpp) by the tests.
package paillier

/*
#include "pgpu.h"
*/
import "C"

import (
	"errors"
	"unsafe"

	gmp "github.com/ncw/gmp"
)

func ctValues(cts []*Ciphertext) []*gmp.Int {
	vals := make([]*gmp.Int, len(cts))
	for i, ct := range cts {
		vals[i] = ct.C
	}
	return vals
}

func wrapCts(vals []*gmp.Int, level EncryptionLevel, method EncryptionMethod) []*Ciphertext {
	out := make([]*Ciphertext, len(vals))
	for i, v := range vals {
		out[i] = &Ciphertext{v, level, method}
	}
	return out
}

// widths of plaintext and ciphertext records at an encryption level (paillier.go:404-414)
func (g *GPUContext) levelWidths(level EncryptionLevel) (int, int, error) {
	switch level {
	case EncLevelOne:
		return g.wN, g.wN2, nil
	case EncLevelTwo:
		return g.wN2, g.wN3, nil
	}
	return 0, 0, errors.New("unsupported encryption level")
}

// PrecomputeRnBatch returns r^n mod n^2 for later EncryptWithRnBatch calls: EncryptWithR(0, r) is exactly that.
func (g *GPUContext) PrecomputeRnBatch(rs []*gmp.Int) ([]*gmp.Int, error) {
	zeros := make([]*gmp.Int, len(rs))
	for i := range zeros {
		zeros[i] = gmp.NewInt(0)
	}
	cts, err := g.EncryptWithRBatch(zeros, rs)
	if err != nil {
		return nil, err
	}
	return ctValues(cts), nil
}

// EncryptWithRnBatch = N x EncryptWithR (paillier.go:206-218) with r^n mod n^2 already computed:
// c = (1 + m*n) * rn mod n^2.  Each rn is one ciphertext's randomness: use it once.
func (g *GPUContext) EncryptWithRnBatch(ms, rns []*gmp.Int) ([]*Ciphertext, error) {
	if len(ms) != len(rns) {
		return nil, errors.New("one r^n per plaintext")
	}
	m, r := toRecords(ms, g.wN), toRecords(rns, g.wN2)
	c := make([]byte, len(ms)*g.wN2)
	if err := gpuErr(g.ctx, C.pgpu_encrypt_with_rn(g.ctx, C.size_t(len(ms)), ptr(m), ptr(r), ptr(c))); err != nil {
		return nil, err
	}
	return wrapCts(fromRecords(c, g.wN2), EncLevelOne, RegularEncryption), nil
}

// EncryptWithRAtLevelBatch = N x PublicKey.EncryptWithRAtLevel (paillier.go:206-218).
func (g *GPUContext) EncryptWithRAtLevelBatch(ms, rs []*gmp.Int, level EncryptionLevel) ([]*Ciphertext, error) {
	if len(ms) != len(rs) {
		return nil, errors.New("one r per plaintext")
	}
	wm, wc, err := g.levelWidths(level)
	if err != nil {
		return nil, err
	}
	m, r := toRecords(ms, wm), toRecords(rs, g.wN)
	c := make([]byte, len(ms)*wc)
	var rc C.int
	if g.secret { // same ciphertexts over the prime powers
		rc = C.pgpu_encrypt_with_r_at_level_sk(g.ctx, C.int(level)+1, C.size_t(len(ms)), ptr(m), ptr(r), ptr(c))
	} else {
		rc = C.pgpu_encrypt_with_r_at_level(g.ctx, C.int(level)+1, C.size_t(len(ms)), ptr(m), ptr(r), ptr(c))
	}
	if err := gpuErr(g.ctx, rc); err != nil {
		return nil, err
	}
	return wrapCts(fromRecords(c, wc), level, RegularEncryption), nil
}

// AltEncryptWithRAtLevelBatch = N x PublicKey.AltEncryptWithRAtLevel (paillier.go:221-238).  Like the scalar
// method (:228) it reduces the caller's r values mod K in place.  The context must come from a key with H, K.
func (g *GPUContext) AltEncryptWithRAtLevelBatch(pk *PublicKey, ms, rs []*gmp.Int, level EncryptionLevel) ([]*Ciphertext, error) {
	if len(ms) != len(rs) {
		return nil, errors.New("one r per plaintext")
	}
	wm, wc, err := g.levelWidths(level)
	if err != nil {
		return nil, err
	}
	for _, r := range rs {
		r.Mod(r, pk.K)
	}
	m, r := toRecords(ms, wm), toRecords(rs, g.wN)
	c := make([]byte, len(ms)*wc)
	if err := gpuErr(g.ctx, C.pgpu_alt_encrypt_with_r_at_level(g.ctx, C.int(level)+1, C.size_t(len(ms)), ptr(m), ptr(r), ptr(c))); err != nil {
		return nil, err
	}
	return wrapCts(fromRecords(c, wc), level, AlternativeEncryption), nil
}

// DecryptAtLevelBatch = N x SecretKey.Decrypt (paillier.go:292-340); all ciphertexts at the same level.
func (g *GPUContext) DecryptAtLevelBatch(cts []*Ciphertext) ([]*gmp.Int, error) {
	if len(cts) == 0 {
		return nil, nil
	}
	level := cts[0].Level
	for _, ct := range cts {
		if ct.Level != level {
			return nil, errors.New("one encryption level per batch")
		}
	}
	wm, wc, err := g.levelWidths(level)
	if err != nil {
		return nil, err
	}
	c := toRecords(ctValues(cts), wc)
	m := make([]byte, len(cts)*wm)
	if err := gpuErr(g.ctx, C.pgpu_decrypt_at_level(g.ctx, C.int(level)+1, C.size_t(len(cts)), ptr(c), ptr(m))); err != nil {
		return nil, err
	}
	return fromRecords(m, wm), nil
}

// NestedDecryptBatch = N x SecretKey.NestedDecrypt (paillier.go:344-356): both layers; an inner value 0 yields 0 (:350-354).
func (g *GPUContext) NestedDecryptBatch(cts []*Ciphertext) ([]*gmp.Int, error) {
	for _, ct := range cts {
		if ct.Level == EncLevelOne {
			panic("no nested ciphertexts to recover") // paillier.go:362-364
		}
	}
	inner, err := g.DecryptAtLevelBatch(cts)
	if err != nil {
		return nil, err
	}
	var nz []*Ciphertext
	var at []int
	for i, v := range inner {
		if v.Sign() != 0 {
			nz = append(nz, &Ciphertext{v, EncLevelOne, MixedEncryption})
			at = append(at, i)
		}
	}
	out := make([]*gmp.Int, len(cts))
	for i := range out {
		out[i] = gmp.NewInt(0)
	}
	vals, err := g.DecryptAtLevelBatch(nz)
	if err != nil {
		return nil, err
	}
	for k, i := range at {
		out[i] = vals[k]
	}
	return out, nil
}

// batchLevel returns the encryption level shared by a batch (operations.go:11-64 take n^(s+1) from the ciphertext's
// level, getModuliForLevel paillier.go:403-414): the record width, the modulus selector of the generic entry points.
func (g *GPUContext) batchLevel(cts []*Ciphertext, what string) (EncryptionLevel, int, C.int, error) {
	level := EncLevelOne
	if len(cts) > 0 {
		level = cts[0].Level
	}
	for _, c := range cts {
		if c.Level != level {
			return level, 0, 0, errors.New(what + ": one encryption level per batch")
		}
	}
	if level == EncLevelOne {
		return level, g.wN2, C.PGPU_MOD_N2, nil
	}
	if g.wN3 == 0 {
		return level, 0, 0, errors.New(what + ": n^3 is wider than the built kernel shapes")
	}
	return level, g.wN3, C.PGPU_MOD_N3, nil
}

// AddPairs = N x PublicKey.Add(a_i, b_i) (operations.go:11-29): modulus and level of a_i, one level per batch.
func (g *GPUContext) AddPairs(as, bs []*Ciphertext) ([]*Ciphertext, error) {
	if len(as) != len(bs) {
		return nil, errors.New("pairs")
	}
	level, w, modsel, err := g.batchLevel(as, "AddPairs")
	if err != nil {
		return nil, err
	}
	a, b := toRecords(ctValues(as), w), toRecords(ctValues(bs), w)
	o := make([]byte, len(as)*w)
	if err := gpuErr(g.ctx, C.pgpu_modmul(g.ctx, modsel, C.size_t(len(as)), ptr(a), ptr(b), ptr(o))); err != nil {
		return nil, err
	}
	return wrapCts(fromRecords(o, w), level, MixedEncryption), nil
}

// SubPairs = N x PublicKey.Sub(a_i, b_i) (operations.go:32-55): modulus and level of a_i, one level per batch.
func (g *GPUContext) SubPairs(as, bs []*Ciphertext) ([]*Ciphertext, error) {
	if len(as) != len(bs) {
		return nil, errors.New("pairs")
	}
	level, w, modsel, err := g.batchLevel(as, "SubPairs")
	if err != nil {
		return nil, err
	}
	a, b := toRecords(ctValues(as), w), toRecords(ctValues(bs), w)
	o := make([]byte, len(as)*w)
	if level == EncLevelOne {
		err = gpuErr(g.ctx, C.pgpu_sub_pairs(g.ctx, C.size_t(len(as)), ptr(a), ptr(b), ptr(o)))
	} else if len(as) > 0 {
		inv := make([]byte, len(b))
		if err = gpuErr(g.ctx, C.pgpu_modinv(g.ctx, modsel, C.size_t(len(as)), ptr(b), ptr(inv))); err == nil {
			err = gpuErr(g.ctx, C.pgpu_modmul(g.ctx, modsel, C.size_t(len(as)), ptr(a), ptr(inv), ptr(o)))
		}
	}
	if err != nil {
		return nil, err
	}
	return wrapCts(fromRecords(o, w), level, MixedEncryption), nil
}

// DotProduct = Add(ConstMult(c_i, k_i)...) with 64-bit scalars in one call (operations.go:11-29,58-64).
func (g *GPUContext) DotProduct(cts []*Ciphertext, ks []uint64) (*Ciphertext, error) {
	if len(cts) != len(ks) {
		return nil, errors.New("one scalar per ciphertext")
	}
	c := toRecords(ctValues(cts), g.wN2)
	o := make([]byte, g.wN2)
	var kp *C.uint64_t
	if len(ks) > 0 {
		kp = (*C.uint64_t)(unsafe.Pointer(&ks[0]))
	}
	if err := gpuErr(g.ctx, C.pgpu_dot_u64(g.ctx, C.size_t(len(cts)), ptr(c), kp, ptr(o))); err != nil {
		return nil, err
	}
	return &Ciphertext{fromRecords(o, g.wN2)[0], EncLevelOne, MixedEncryption}, nil
}

// RandomizeWithRBatch = N x PublicKey.Randomize (operations.go:67-69) with the r of the fresh Encrypt(0) supplied.
func (g *GPUContext) RandomizeWithRBatch(cts []*Ciphertext, rs []*gmp.Int) ([]*Ciphertext, error) {
	if len(cts) != len(rs) {
		return nil, errors.New("one r per ciphertext")
	}
	c, r := toRecords(ctValues(cts), g.wN2), toRecords(rs, g.wN)
	o := make([]byte, len(cts)*g.wN2)
	if err := gpuErr(g.ctx, C.pgpu_randomize_with_r(g.ctx, C.size_t(len(cts)), ptr(c), ptr(r), ptr(o))); err != nil {
		return nil, err
	}
	return wrapCts(fromRecords(o, g.wN2), EncLevelOne, MixedEncryption), nil
}

// ExtractRandonnessBatch = N x SecretKey.ExtractRandonness (operations.go:75-91); one level per batch.
func (g *GPUContext) ExtractRandonnessBatch(cts []*Ciphertext) ([]*gmp.Int, error) {
	if len(cts) == 0 {
		return nil, nil
	}
	_, wc, err := g.levelWidths(cts[0].Level)
	if err != nil {
		return nil, err
	}
	c := toRecords(ctValues(cts), wc)
	o := make([]byte, len(cts)*g.wN)
	if err := gpuErr(g.ctx, C.pgpu_extract_randomness(g.ctx, C.int(cts[0].Level)+1, C.size_t(len(cts)), ptr(c), ptr(o))); err != nil {
		return nil, err
	}
	return fromRecords(o, g.wN), nil
}

// NestedRandomizeWithBatch = N x PublicKey.NestedRandomize (operations.go:96-118) with (a, b) supplied.
func (g *GPUContext) NestedRandomizeWithBatch(cts []*Ciphertext, as, bs []*gmp.Int) ([]*Ciphertext, error) {
	for _, ct := range cts {
		if ct.Level != EncLevelTwo {
			panic("can only homomorphically randomize doubly encrypted values") // operations.go:97-99
		}
	}
	c, a, b := toRecords(ctValues(cts), g.wN3), toRecords(as, g.wN), toRecords(bs, g.wN)
	o := make([]byte, len(cts)*g.wN3)
	if err := gpuErr(g.ctx, C.pgpu_nested_randomize_with(g.ctx, C.size_t(len(cts)), ptr(c), ptr(a), ptr(b), ptr(o))); err != nil {
		return nil, err
	}
	return wrapCts(fromRecords(o, g.wN3), EncLevelTwo, RegularEncryption), nil
}

func (g *GPUContext) nested(sub bool, ct1, ct2 []*Ciphertext) ([]*Ciphertext, error) {
	if len(ct1) != len(ct2) {
		return nil, errors.New("pairs")
	}
	for i := range ct1 {
		if ct1[i].Level != EncLevelTwo || ct2[i].Level != EncLevelOne {
			panic("can only homomorphically add an encrypted value to a doubly encrypted value") // operations.go:122-124
		}
	}
	a, b := toRecords(ctValues(ct1), g.wN3), toRecords(ctValues(ct2), g.wN2)
	o := make([]byte, len(ct1)*g.wN3)
	var rc C.int
	if sub {
		rc = C.pgpu_nested_sub(g.ctx, C.size_t(len(ct1)), ptr(a), ptr(b), ptr(o))
	} else {
		rc = C.pgpu_nested_add(g.ctx, C.size_t(len(ct1)), ptr(a), ptr(b), ptr(o))
	}
	if err := gpuErr(g.ctx, rc); err != nil {
		return nil, err
	}
	out := make([]*Ciphertext, len(ct1))
	for i, v := range fromRecords(o, g.wN3) {
		out[i] = &Ciphertext{v, ct1[i].Level, ct1[i].EncMethod}
	}
	return out, nil
}

// NestedAddBatch / NestedSubBatch = N x PublicKey.NestedAdd / NestedSub (operations.go:121-140).
func (g *GPUContext) NestedAddBatch(ct1, ct2 []*Ciphertext) ([]*Ciphertext, error) { return g.nested(false, ct1, ct2) }
func (g *GPUContext) NestedSubBatch(ct1, ct2 []*Ciphertext) ([]*Ciphertext, error) { return g.nested(true, ct1, ct2) }

// VerifyProofBatch = N x PartialDecryptionZKP.VerifyProof (thresholdkey.go:278-311) for proofs of one server.
func (g *GPUContext) VerifyProofBatch(proofs []*PartialDecryptionZKP) ([]bool, error) {
	if len(proofs) == 0 {
		return nil, nil
	}
	id := proofs[0].ID
	cs, ds, es, zs := make([]*gmp.Int, len(proofs)), make([]*gmp.Int, len(proofs)), make([]*gmp.Int, len(proofs)), make([]*gmp.Int, len(proofs))
	fits := make([]bool, len(proofs))
	zero := gmp.NewInt(0)
	for i, p := range proofs {
		if p.ID != id {
			return nil, errors.New("one server id per batch")
		}
		// a value wider than its record cannot come from an honest prover (E is a SHA-256 digest, Z < 2^(8 wZ) for every
		// r < n^2): the scalar VerifyProof answers false for it, so does the batch instead of failing as a whole
		fits[i] = p.C.Sign() >= 0 && p.Decryption.Sign() >= 0 && p.E.Sign() >= 0 && p.Z.Sign() >= 0 &&
			len(p.C.Bytes()) <= g.wN2 && len(p.Decryption.Bytes()) <= g.wN2 && len(p.E.Bytes()) <= 32 && len(p.Z.Bytes()) <= g.wZ
		if fits[i] {
			cs[i], ds[i], es[i], zs[i] = p.C, p.Decryption, p.E, p.Z
		} else {
			cs[i], ds[i], es[i], zs[i] = zero, zero, zero, zero
		}
	}
	c, d, e, z := toRecords(cs, g.wN2), toRecords(ds, g.wN2), toRecords(es, 32), toRecords(zs, g.wZ)
	ok := make([]byte, len(proofs))
	rc := C.pgpu_pdec_zkp_verify(g.ctx, C.size_t(len(proofs)), C.int(id), ptr(c), ptr(d), ptr(e), ptr(z), (*C.uint8_t)(ptr(ok)))
	if err := gpuErr(g.ctx, rc); err != nil {
		return nil, err
	}
	out := make([]bool, len(proofs))
	for i, v := range ok {
		out[i] = v == 1 && fits[i]
	}
	return out, nil
}

// VerifyDDLEQProofBatch = N x PublicKey.VerifyDDLEQProof (ddleq.go:44-53); one instance count per batch.
func (g *GPUContext) VerifyDDLEQProofBatch(ct1, ct2 []*Ciphertext, proofs []*DDLEQProof) ([]bool, error) {
	out := make([]bool, len(proofs))
	if len(proofs) == 0 {
		return out, nil
	}
	secpar := len(proofs[0].Instances)
	for i := range out {
		out[i] = true // an empty instance list verifies (ddleq.go:46-51)
	}
	if secpar == 0 {
		return out, nil
	}
	var xs, ys, as, es, fs []*gmp.Int
	for _, p := range proofs {
		if len(p.Instances) != secpar {
			return nil, errors.New("one secpar per batch")
		}
		for _, in := range p.Instances {
			xs, ys, as, es, fs = append(xs, in.X), append(ys, in.Y), append(as, in.Alpha), append(es, in.E), append(fs, in.F)
		}
	}
	c1, c2 := toRecords(ctValues(ct1), g.wN3), toRecords(ctValues(ct2), g.wN3)
	x, y, al, e, f := toRecords(xs, g.wN), toRecords(ys, g.wN), toRecords(as, g.wN3), toRecords(es, g.wN2), toRecords(fs, g.wN3)
	ok := make([]byte, len(xs))
	rc := C.pgpu_ddleq_verify(g.ctx, C.size_t(len(proofs)), C.uint(secpar), ptr(c1), ptr(c2), ptr(x), ptr(y), ptr(al), ptr(e), ptr(f), (*C.uint8_t)(ptr(ok)))
	if err := gpuErr(g.ctx, rc); err != nil {
		return nil, err
	}
	for k, v := range ok {
		if v != 1 {
			out[k/secpar] = false
		}
	}
	return out, nil
}

// ExpBatch = N x gmp.Int.Exp(base_i, exp_i, modulus) against n (modsel 0), n^2 (1) or n^3 (2): the batched form of
// createVerificationKeys' loop (thresholdkey_generator.go:246-254) and of any other Exp of the package.
func (g *GPUContext) ExpBatch(modsel int, bases, exps []*gmp.Int) ([]*gmp.Int, error) {
	if len(bases) != len(exps) {
		return nil, errors.New("one exponent per base")
	}
	w := []int{g.wN, g.wN2, g.wN3}[modsel]
	eBytes := 4
	for _, e := range exps {
		if n := (len(e.Bytes()) + 3) / 4 * 4; n > eBytes {
			eBytes = n
		}
	}
	b, e := toRecords(bases, w), toRecords(exps, eBytes)
	o := make([]byte, len(bases)*w)
	if err := gpuErr(g.ctx, C.pgpu_modexp(g.ctx, C.int(modsel), C.size_t(len(bases)), ptr(b), ptr(e), C.size_t(eBytes), ptr(o))); err != nil {
		return nil, err
	}
	return fromRecords(o, w), nil
}
