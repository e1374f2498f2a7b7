// batch_gpu.go -- the file a maintainer of sachaservan/paillier adds next to paillier.go to bind
// libpaillier_b200.so.  It adds BATCH methods beside the scalar API (which stays untouched); every
// method is marshalling only: gmp.Int -> fixed-width little-endian records, one cgo call, records ->
// gmp.Int.  NOT COMPILED in this repository's image (no Go toolchain); the same symbols are exercised
// through ctypes by paillier_b200/api.py and the tests.
//
// Build: CGO_CFLAGS="-I${PAILLIER_B200}/include" CGO_LDFLAGS="-L${PAILLIER_B200}/paillier_b200 -lpaillier_b200" go build
package paillier

/*
#cgo LDFLAGS: -lpaillier_b200
#include <stdlib.h>
#include "pgpu.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"runtime"
	"unsafe"

	gmp "github.com/ncw/gmp"
)

// GPUContext owns one engine context: one key on one device (pgpu_ctx_create).
type GPUContext struct {
	ctx            *C.pgpu_ctx
	wN, wN2, wN3   int
	wZ             int
	secret         bool // p, q are loaded: EncryptWithRBatch may use them
}

func gpuErr(ctx *C.pgpu_ctx, rc C.int) error {
	if rc == C.PGPU_OK {
		return nil
	}
	return fmt.Errorf("paillier_b200: error %d: %s", int(rc), C.GoString(C.pgpu_last_error(ctx)))
}

// NewGPUContext binds pk to `device` (PublicKey, paillier.go:46-56).
func (pk *PublicKey) NewGPUContext(device int) (*GPUContext, error) {
	nb := pk.N.Bytes()
	g := &GPUContext{}
	if err := gpuErr(nil, C.pgpu_ctx_create(&g.ctx, C.int(device), (*C.uint8_t)(unsafe.Pointer(&nb[0])), C.size_t(len(nb)))); err != nil {
		return nil, err
	}
	var a, b, c C.size_t
	C.pgpu_ctx_widths(g.ctx, &a, &b, &c)
	g.wN, g.wN2, g.wN3 = int(a), int(b), int(c)
	if pk.H != nil && pk.K != nil {
		hb := pk.H.Bytes()
		if err := gpuErr(g.ctx, C.pgpu_ctx_set_alt_generator(g.ctx, (*C.uint8_t)(unsafe.Pointer(&hb[0])), C.size_t(len(hb)), C.uint(pk.K.BitLen()-1))); err != nil {
			g.Close()
			return nil, err
		}
	}
	runtime.SetFinalizer(g, func(x *GPUContext) { x.Close() })
	return g, nil
}

// NewGPUContext for a SecretKey also loads Lambda (paillier.go:59-62); the engine recovers p, q.
func (sk *SecretKey) NewGPUContext(device int) (*GPUContext, error) {
	g, err := sk.PublicKey.NewGPUContext(device)
	if err != nil {
		return nil, err
	}
	lb := sk.Lambda.Bytes()
	if err := gpuErr(g.ctx, C.pgpu_ctx_set_secret_lambda(g.ctx, (*C.uint8_t)(unsafe.Pointer(&lb[0])), C.size_t(len(lb)))); err != nil {
		g.Close()
		return nil, err
	}
	g.secret = true
	return g, nil
}

func (g *GPUContext) Close() {
	if g.ctx != nil {
		C.pgpu_ctx_destroy(g.ctx)
		g.ctx = nil
	}
}

// toRecords writes each value as a `width`-byte little-endian record (gmp.Int.Bytes() is big-endian).
func toRecords(vals []*gmp.Int, width int) []byte {
	buf := make([]byte, len(vals)*width)
	for i, v := range vals {
		be := v.Bytes()
		if len(be) > width {
			panic("paillier_b200: value wider than its record")
		}
		rec := buf[i*width : (i+1)*width]
		for j, b := range be {
			rec[len(be)-1-j] = b
		}
	}
	return buf
}

func fromRecords(buf []byte, width int) []*gmp.Int {
	out := make([]*gmp.Int, len(buf)/width)
	be := make([]byte, width)
	for i := range out {
		rec := buf[i*width : (i+1)*width]
		for j := 0; j < width; j++ {
			be[width-1-j] = rec[j]
		}
		out[i] = new(gmp.Int).SetBytes(be)
	}
	return out
}

func ptr(b []byte) unsafe.Pointer {
	if len(b) == 0 {
		return nil
	}
	return unsafe.Pointer(&b[0])
}

// EncryptWithRBatch = N x PublicKey.EncryptWithR (paillier.go:185-187).
func (g *GPUContext) EncryptWithRBatch(ms, rs []*gmp.Int) ([]*Ciphertext, error) {
	if len(ms) != len(rs) {
		return nil, errors.New("one r per plaintext")
	}
	m, r := toRecords(ms, g.wN), toRecords(rs, g.wN)
	c := make([]byte, len(ms)*g.wN2)
	var rc C.int
	if g.secret { // context made from a SecretKey: r^n over p^2 and q^2, the same ciphertexts at about twice the rate
		rc = C.pgpu_encrypt_with_r_sk(g.ctx, C.size_t(len(ms)), ptr(m), ptr(r), ptr(c))
	} else {
		rc = C.pgpu_encrypt_with_r(g.ctx, C.size_t(len(ms)), ptr(m), ptr(r), ptr(c))
	}
	if err := gpuErr(g.ctx, rc); err != nil {
		return nil, err
	}
	out := make([]*Ciphertext, len(ms))
	for i, v := range fromRecords(c, g.wN2) {
		out[i] = &Ciphertext{v, EncLevelOne, RegularEncryption}
	}
	return out, nil
}

// DecryptBatch = N x SecretKey.Decrypt (paillier.go:292-303), level 1.
func (g *GPUContext) DecryptBatch(cts []*Ciphertext) ([]*gmp.Int, error) {
	vals := make([]*gmp.Int, len(cts))
	for i, ct := range cts {
		vals[i] = ct.C
	}
	c := toRecords(vals, g.wN2)
	m := make([]byte, len(cts)*g.wN)
	if err := gpuErr(g.ctx, C.pgpu_decrypt(g.ctx, C.size_t(len(cts)), ptr(c), ptr(m))); err != nil {
		return nil, err
	}
	return fromRecords(m, g.wN), nil
}

// ConstMultBatch = N x PublicKey.ConstMult (operations.go:58-64); k <= 0 yields 1 like gmp's Exp.
func (g *GPUContext) ConstMultBatch(cts []*Ciphertext, ks []*gmp.Int) ([]*Ciphertext, error) {
	kBytes := 4
	zero := gmp.NewInt(0)
	kk := make([]*gmp.Int, len(ks))
	for i, k := range ks {
		kk[i] = k
		if k.Sign() <= 0 {
			kk[i] = zero
		}
		if n := (len(kk[i].Bytes()) + 3) / 4 * 4; n > kBytes {
			kBytes = n
		}
	}
	vals := make([]*gmp.Int, len(cts))
	for i, ct := range cts {
		vals[i] = ct.C
	}
	_, w, modsel, err := g.batchLevel(cts, "ConstMultBatch")
	if err != nil {
		return nil, err
	}
	c, k := toRecords(vals, w), toRecords(kk, kBytes)
	o := make([]byte, len(cts)*w)
	if err := gpuErr(g.ctx, C.pgpu_modexp(g.ctx, modsel, C.size_t(len(cts)), ptr(c), ptr(k), C.size_t(kBytes), ptr(o))); err != nil {
		return nil, err
	}
	out := make([]*Ciphertext, len(cts))
	for i, v := range fromRecords(o, w) {
		out[i] = &Ciphertext{v, cts[i].Level, cts[i].EncMethod}
	}
	return out, nil
}

// AddBatch = PublicKey.Add(cts...) (operations.go:11-29) as one tree reduction; modulus and level of cts[0].
func (g *GPUContext) AddBatch(cts []*Ciphertext) (*Ciphertext, error) {
	vals := make([]*gmp.Int, len(cts))
	for i, ct := range cts {
		vals[i] = ct.C
	}
	level, w, lv := EncLevelOne, g.wN2, C.int(1)
	if len(cts) > 0 && cts[0].Level == EncLevelTwo {
		level, w, lv = EncLevelTwo, g.wN3, C.int(2)
	}
	c := toRecords(vals, w)
	o := make([]byte, w)
	if err := gpuErr(g.ctx, C.pgpu_add_reduce_at_level(g.ctx, lv, C.size_t(len(cts)), ptr(c), ptr(o))); err != nil {
		return nil, err
	}
	return &Ciphertext{fromRecords(o, w)[0], level, MixedEncryption}, nil
}

// NewGPUContext for a threshold key share (thresholdkey.go:26-42).
func (tsk *ThresholdSecretKey) NewGPUContext(device int) (*GPUContext, error) {
	pk := &PublicKey{N: tsk.N}
	g, err := pk.NewGPUContext(device)
	if err != nil {
		return nil, err
	}
	sb, vb := tsk.Share.Bytes(), tsk.VerificationKey.Bytes()
	vk := toRecords(tsk.VerificationKeys, g.wN2)
	rc := C.pgpu_ctx_set_threshold(g.ctx, C.int(tsk.TotalNumberOfDecryptionServers), C.int(tsk.Threshold), C.int(tsk.ID),
		(*C.uint8_t)(unsafe.Pointer(&sb[0])), C.size_t(len(sb)), (*C.uint8_t)(unsafe.Pointer(&vb[0])), C.size_t(len(vb)), ptr(vk))
	if err := gpuErr(g.ctx, rc); err != nil {
		g.Close()
		return nil, err
	}
	var wz C.size_t
	C.pgpu_ctx_z_width(g.ctx, &wz)
	g.wZ = int(wz)
	return g, nil
}

// PartialDecryptBatch = N x ThresholdSecretKey.PartialDecrypt (thresholdkey.go:192-201).
func (g *GPUContext) PartialDecryptBatch(id int, cs []*gmp.Int) ([]*PartialDecryption, error) {
	c := toRecords(cs, g.wN2)
	o := make([]byte, len(cs)*g.wN2)
	if err := gpuErr(g.ctx, C.pgpu_partial_decrypt(g.ctx, C.size_t(len(cs)), ptr(c), ptr(o))); err != nil {
		return nil, err
	}
	out := make([]*PartialDecryption, len(cs))
	for i, v := range fromRecords(o, g.wN2) {
		out[i] = &PartialDecryption{ID: id, Decryption: v}
	}
	return out, nil
}

// PartialDecryptionWithZKPBatch = N x PartialDecryptionWithZKP (thresholdkey.go:225-255); rs are the
// r in [0, n^2) the scalar method draws from crypto/rand at :233.
func (g *GPUContext) PartialDecryptionWithZKPBatch(tk *ThresholdPublicKey, id int, cs, rs []*gmp.Int) ([]*PartialDecryptionZKP, error) {
	c, r := toRecords(cs, g.wN2), toRecords(rs, g.wN2)
	dec, e, z := make([]byte, len(cs)*g.wN2), make([]byte, len(cs)*32), make([]byte, len(cs)*g.wZ)
	if err := gpuErr(g.ctx, C.pgpu_pdec_zkp_prove(g.ctx, C.size_t(len(cs)), ptr(c), ptr(r), ptr(dec), ptr(e), ptr(z))); err != nil {
		return nil, err
	}
	D, E, Z := fromRecords(dec, g.wN2), fromRecords(e, 32), fromRecords(z, g.wZ)
	out := make([]*PartialDecryptionZKP, len(cs))
	for i := range cs {
		out[i] = &PartialDecryptionZKP{PartialDecryption: PartialDecryption{ID: id, Decryption: D[i]}, Key: tk, E: E[i], Z: Z[i], C: cs[i]}
	}
	return out, nil
}

// CombinePartialDecryptionsBatch = N x CombinePartialDecryptions (thresholdkey.go:149-161); shares[j]
// is server j's batch, all batches in the same ciphertext order.
func (g *GPUContext) CombinePartialDecryptionsBatch(shares [][]*PartialDecryption) ([]*gmp.Int, error) {
	k := len(shares)
	if k == 0 {
		return nil, errors.New("Threshold not meet")
	}
	count := len(shares[0])
	ids := make([]C.int, k)
	flat := make([]*gmp.Int, 0, k*count)
	for j, s := range shares {
		ids[j] = C.int(s[0].ID)
		for _, pd := range s {
			flat = append(flat, pd.Decryption)
		}
	}
	d := toRecords(flat, g.wN2)
	m := make([]byte, count*g.wN)
	rc := C.pgpu_combine(g.ctx, C.size_t(count), C.int(k), &ids[0], ptr(d), ptr(m))
	if rc == C.PGPU_ERR_THRESHOLD {
		return nil, errors.New(C.GoString(C.pgpu_last_error(g.ctx))) // "Threshold not meet" / duplicate server, thresholdkey.go:77-89
	}
	if err := gpuErr(g.ctx, rc); err != nil {
		return nil, err
	}
	return fromRecords(m, g.wN), nil
}

// ProveDDLEQBatch = N x SecretKey.ProveDDLEQ (ddleq.go:27-40) with the per-instance x, y supplied.
// Panics where the scalar method panics (ddleq.go:67-69).
func (g *GPUContext) ProveDDLEQBatch(secpar int, ct1, ct2 []*Ciphertext, as, bs []*gmp.Int, xs, ys [][]*gmp.Int) ([]*DDLEQProof, error) {
	count := len(ct1)
	c1v, c2v := make([]*gmp.Int, count), make([]*gmp.Int, count)
	var fx, fy []*gmp.Int
	for i := range ct1 {
		c1v[i], c2v[i] = ct1[i].C, ct2[i].C
		fx, fy = append(fx, xs[i]...), append(fy, ys[i]...)
	}
	total := count * secpar
	al, e, f := make([]byte, total*g.wN3), make([]byte, total*g.wN2), make([]byte, total*g.wN3)
	c1, c2, a, b, x, y := toRecords(c1v, g.wN3), toRecords(c2v, g.wN3), toRecords(as, g.wN), toRecords(bs, g.wN), toRecords(fx, g.wN), toRecords(fy, g.wN)
	rc := C.pgpu_ddleq_prove(g.ctx, C.size_t(count), C.uint(secpar), ptr(c1), ptr(c2), ptr(a), ptr(b), ptr(x), ptr(y), ptr(al), ptr(e), ptr(f))
	if rc == C.PGPU_ERR_ARG {
		panic("cannot prove re-encryption because inputs are wrong")
	}
	if err := gpuErr(g.ctx, rc); err != nil {
		return nil, err
	}
	A, E, F := fromRecords(al, g.wN3), fromRecords(e, g.wN2), fromRecords(f, g.wN3)
	out := make([]*DDLEQProof, count)
	for i := range out {
		p := &DDLEQProof{Instances: make([]*DDLEQProofInstance, secpar)}
		for j := 0; j < secpar; j++ {
			k := i*secpar + j
			p.Instances[j] = &DDLEQProofInstance{X: fx[k], Y: fy[k], Alpha: A[k], E: E[k], F: F[k]}
		}
		out[i] = p
	}
	return out, nil
}

// GenerateSafePrimeGPU = GenerateSafePrime (safe_prime.go:61-105) with the candidate loop on the GPU:
// reads `batch` candidates at a time from `random` and returns the first accepted one in stream order.
func GenerateSafePrimeGPU(bitLen, device, batch, maxBatches int, random interface{ Read([]byte) (int, error) }) (*gmp.Int, *gmp.Int, error) {
	if bitLen < 6 {
		return nil, nil, errors.New("safe prime size must be at least 6 bits")
	}
	nb := (bitLen - 1 + 7) / 8
	s := 32
	if bitLen > 1024 {
		s = 48
	}
	if bitLen > 1536 {
		s = 64
	}
	raw := make([]byte, batch*nb)
	p, q, ok := make([]byte, batch*s*4), make([]byte, batch*s*4), make([]byte, batch)
	for it := 0; it < maxBatches; it++ {
		if _, err := random.Read(raw); err != nil {
			return nil, nil, err
		}
		rc := C.pgpu_safe_prime_scan(C.int(device), C.uint(bitLen), C.size_t(batch), (*C.uint8_t)(ptr(raw)), ptr(p), ptr(q), (*C.uint8_t)(ptr(ok)), nil)
		if rc != C.PGPU_OK {
			return nil, nil, fmt.Errorf("paillier_b200: %s", C.GoString(C.pgpu_primes_last_error()))
		}
		for i := 0; i < batch; i++ {
			if ok[i] == 1 {
				return fromRecords(p[i*s*4:(i+1)*s*4], s*4)[0], fromRecords(q[i*s*4:(i+1)*s*4], s*4)[0], nil
			}
		}
	}
	return nil, nil, fmt.Errorf("generator gave up after %d batches", maxBatches)
}
