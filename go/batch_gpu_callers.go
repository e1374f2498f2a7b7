// batch_gpu_callers.go -- batch forms of the reference's callers around the hot path: the methods that draw their own
// randomness (paillier.go:192-203,244-289, operations.go:67-69,96-118, thresholdkey.go:258-275) and the variadic Sub
// (operations.go:32-55).  The draws stay on the host and use the package's own GetRandomNumberInMultiplicativeGroup
// (utils.go:36-49, crypto/rand); everything per item runs in libpaillier_b200.so.  NOT COMPILED in this repository's
// image (no Go toolchain); paillier_b200/api.py and include/paillier_b200.hpp carry the same methods and are tested.
package paillier

/*
#include "pgpu.h"
*/
import "C"

import (
	"crypto/rand"
	"errors"

	gmp "github.com/ncw/gmp"
)

// drawUnits = count x GetRandomNumberInMultiplicativeGroup(n, rand.Reader) (utils.go:36-49).
func drawUnits(n *gmp.Int, count int) ([]*gmp.Int, error) {
	rs := make([]*gmp.Int, count)
	for i := range rs {
		r, err := GetRandomNumberInMultiplicativeGroup(n, rand.Reader)
		if err != nil {
			return nil, err
		}
		rs[i] = r
	}
	return rs, nil
}

// EncryptAtLevelBatch = N x PublicKey.EncryptAtLevel (paillier.go:258-269).
func (g *GPUContext) EncryptAtLevelBatch(pk *PublicKey, ms []*gmp.Int, level EncryptionLevel) ([]*Ciphertext, error) {
	rs, err := drawUnits(pk.N, len(ms))
	if err != nil {
		return nil, err
	}
	return g.EncryptWithRAtLevelBatch(ms, rs, level)
}

// EncryptBatch = N x PublicKey.Encrypt (paillier.go:192-194).
func (g *GPUContext) EncryptBatch(pk *PublicKey, ms []*gmp.Int) ([]*Ciphertext, error) {
	return g.EncryptAtLevelBatch(pk, ms, DefaultEncryptionLevel)
}

// NestedEncryptBatch = N x PublicKey.NestedEncrypt (paillier.go:200-203).
func (g *GPUContext) NestedEncryptBatch(pk *PublicKey, ms []*gmp.Int) ([]*Ciphertext, error) {
	inner, err := g.EncryptAtLevelBatch(pk, ms, EncLevelOne)
	if err != nil {
		return nil, err
	}
	return g.EncryptAtLevelBatch(pk, ctValues(inner), EncLevelTwo)
}

// AltEncryptAtLevelBatch = N x PublicKey.AltEncryptAtLevel (paillier.go:244-255).
func (g *GPUContext) AltEncryptAtLevelBatch(pk *PublicKey, ms []*gmp.Int, level EncryptionLevel) ([]*Ciphertext, error) {
	rs, err := drawUnits(pk.N, len(ms))
	if err != nil {
		return nil, err
	}
	return g.AltEncryptWithRAtLevelBatch(pk, ms, rs, level)
}

func repeated(v int64, count int) []*gmp.Int {
	out := make([]*gmp.Int, count)
	for i := range out {
		out[i] = gmp.NewInt(v)
	}
	return out
}

// EncryptZeroAtLevelBatch / EncryptOneAtLevelBatch = count x EncryptZeroAtLevel / EncryptOneAtLevel (paillier.go:282-289).
func (g *GPUContext) EncryptZeroAtLevelBatch(pk *PublicKey, count int, level EncryptionLevel) ([]*Ciphertext, error) {
	return g.EncryptAtLevelBatch(pk, repeated(0, count), level)
}

func (g *GPUContext) EncryptOneAtLevelBatch(pk *PublicKey, count int, level EncryptionLevel) ([]*Ciphertext, error) {
	return g.EncryptAtLevelBatch(pk, repeated(1, count), level)
}

// EncryptZeroBatch / EncryptOneBatch = count x EncryptZero / EncryptOne (paillier.go:272-279).
func (g *GPUContext) EncryptZeroBatch(pk *PublicKey, count int) ([]*Ciphertext, error) {
	return g.EncryptZeroAtLevelBatch(pk, count, DefaultEncryptionLevel)
}

func (g *GPUContext) EncryptOneBatch(pk *PublicKey, count int) ([]*Ciphertext, error) {
	return g.EncryptOneAtLevelBatch(pk, count, DefaultEncryptionLevel)
}

// RandomizeBatch = N x PublicKey.Randomize (operations.go:67-69) with the r of each fresh Encrypt(0) drawn here.  Add takes
// the modulus from ct.Level while Encrypt(0) is a level-1 ciphertext: a level-2 ct is multiplied by r^n mod n^2 modulo n^3
// (bit-exact with the scalar method; that product is not an encryption of the same plaintext -- use NestedRandomizeBatch).
func (g *GPUContext) RandomizeBatch(pk *PublicKey, cts []*Ciphertext) ([]*Ciphertext, error) {
	rs, err := drawUnits(pk.N, len(cts))
	if err != nil {
		return nil, err
	}
	level, w, modsel, err := g.batchLevel(cts, "RandomizeBatch")
	if err != nil {
		return nil, err
	}
	if level == EncLevelOne {
		return g.RandomizeWithRBatch(cts, rs)
	}
	zeros, err := g.EncryptWithRBatch(repeated(0, len(cts)), rs)
	if err != nil {
		return nil, err
	}
	a, b := toRecords(ctValues(cts), w), toRecords(ctValues(zeros), w)
	o := make([]byte, len(cts)*w)
	if err := gpuErr(g.ctx, C.pgpu_modmul(g.ctx, modsel, C.size_t(len(cts)), ptr(a), ptr(b), ptr(o))); err != nil {
		return nil, err
	}
	return wrapCts(fromRecords(o, w), level, MixedEncryption), nil
}

// NestedRandomizeBatch = N x PublicKey.NestedRandomize (operations.go:96-118): a, b drawn from Z*_n (:105-106) and
// returned with the randomized ciphertexts, like there.
func (g *GPUContext) NestedRandomizeBatch(pk *PublicKey, cts []*Ciphertext) ([]*Ciphertext, []*gmp.Int, []*gmp.Int, error) {
	as, err := drawUnits(pk.N, len(cts))
	if err != nil {
		return nil, nil, nil, err
	}
	bs, err := drawUnits(pk.N, len(cts))
	if err != nil {
		return nil, nil, nil, err
	}
	out, err := g.NestedRandomizeWithBatch(cts, as, bs)
	return out, as, bs, err
}

// SubBatch = PublicKey.Sub(cts...) (operations.go:32-55): cts[0] * prod_{i>0} cts[i]^-1 modulo n^(s+1) of cts[0].Level.
// The inverses of the scalar loop are taken once, of the product of cts[1:] (the same canonical residue); a single
// argument comes back as it is (:34).
func (g *GPUContext) SubBatch(cts []*Ciphertext) (*Ciphertext, error) {
	if len(cts) == 0 {
		return nil, errors.New("Sub needs at least one ciphertext")
	}
	level := cts[0].Level
	if len(cts) == 1 {
		return &Ciphertext{cts[0].C, level, MixedEncryption}, nil
	}
	w, modsel, lv := g.wN2, C.int(C.PGPU_MOD_N2), C.int(1)
	if level == EncLevelTwo {
		if g.wN3 == 0 {
			return nil, errors.New("SubBatch: n^3 is wider than the built kernel shapes")
		}
		w, modsel, lv = g.wN3, C.int(C.PGPU_MOD_N3), C.int(2)
	}
	first, rest := toRecords(ctValues(cts[:1]), w), toRecords(ctValues(cts[1:]), w)
	prod, inv, o := make([]byte, w), make([]byte, w), make([]byte, w)
	if err := gpuErr(g.ctx, C.pgpu_add_reduce_at_level(g.ctx, lv, C.size_t(len(cts)-1), ptr(rest), ptr(prod))); err != nil {
		return nil, err
	}
	if err := gpuErr(g.ctx, C.pgpu_modinv(g.ctx, modsel, 1, ptr(prod), ptr(inv))); err != nil {
		return nil, err
	}
	if err := gpuErr(g.ctx, C.pgpu_modmul(g.ctx, modsel, 1, ptr(first), ptr(inv), ptr(o))); err != nil {
		return nil, err
	}
	return &Ciphertext{fromRecords(o, w)[0], level, MixedEncryption}, nil
}

// VerifyPartialDecryptionBatch = count x ThresholdSecretKey.VerifyPartialDecryption (thresholdkey.go:258-275) in one
// batch: encrypt random m < n, prove the partial decryptions (r < n^2, :233), verify the proofs.  g must come from
// tsk.NewGPUContext.
func (g *GPUContext) VerifyPartialDecryptionBatch(tsk *ThresholdSecretKey, count int) error {
	ms, rs := make([]*gmp.Int, count), make([]*gmp.Int, count)
	for i := range ms {
		m, err := GetRandomNumber(tsk.N, rand.Reader)
		if err != nil {
			return err
		}
		r, err := GetRandomNumber(tsk.GetN2(), rand.Reader)
		if err != nil {
			return err
		}
		ms[i], rs[i] = m, r
	}
	cts, err := g.EncryptBatch(&tsk.ThresholdPublicKey.PublicKey, ms) // tsk.PublicKey is the method of thresholdkey.go:213
	if err != nil {
		return err
	}
	proofs, err := g.PartialDecryptionWithZKPBatch(tsk.PublicKey(), tsk.ID, ctValues(cts), rs)
	if err != nil {
		return err
	}
	oks, err := g.VerifyProofBatch(proofs)
	if err != nil {
		return err
	}
	for _, ok := range oks {
		if !ok {
			return errors.New("Invalid share")
		}
	}
	return nil
}
