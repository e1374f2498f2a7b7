// batch_gpu_device.go -- device-resident ciphertexts, pinned host memory, per-ciphertext proof filtering and the
// several-GPU threshold round (SURVEY.md 8f rank 1, 8b "Ownership" / "Threading", 8e).  Marshalling only, like
// batch_gpu.go; NOT COMPILED in this repository's image (no Go toolchain).  The same symbols are driven through ctypes
// by tests/test_gpu_buffers.py and tests/test_gpu_multi.py.
package paillier

/*
#include <stdlib.h>
#include "pgpu.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"unsafe"

	gmp "github.com/ncw/gmp"
)

// DeviceBuffer is device memory owned by a GPUContext: ciphertext records stay in it between calls, so that
// Encrypt -> ConstMult -> Add -> Decrypt (the callers of operations.go:11-64) never cross PCIe in between.
type DeviceBuffer struct {
	buf *C.pgpu_buf
	g   *GPUContext
}

// NewDeviceBuffer allocates `bytes` on the context's device (pgpu_buf_alloc).
func (g *GPUContext) NewDeviceBuffer(bytes int) (*DeviceBuffer, error) {
	d := &DeviceBuffer{g: g}
	if err := gpuErr(g.ctx, C.pgpu_buf_alloc(g.ctx, C.size_t(bytes), &d.buf)); err != nil {
		return nil, err
	}
	return d, nil
}

// Free waits for the work enqueued on the context, then releases the memory.
func (d *DeviceBuffer) Free() {
	if d.buf != nil {
		C.pgpu_buf_free(d.buf)
		d.buf = nil
	}
}

// Len is the size in bytes.
func (d *DeviceBuffer) Len() int { return int(C.pgpu_buf_size(d.buf)) }

func (d *DeviceBuffer) ptr() unsafe.Pointer { return C.pgpu_buf_ptr(d.buf) }

// Upload copies records into the buffer at byte offset `off`; blocks until done (no Go pointer is retained).
func (d *DeviceBuffer) Upload(off int, records []byte) error {
	return gpuErr(d.g.ctx, C.pgpu_buf_upload(d.buf, C.size_t(off), ptr(records), C.size_t(len(records))))
}

// Download copies len(records) bytes at byte offset `off` back; blocks until done.
func (d *DeviceBuffer) Download(off int, records []byte) error {
	return gpuErr(d.g.ctx, C.pgpu_buf_download(d.buf, C.size_t(off), ptr(records), C.size_t(len(records))))
}

// Sync waits for everything the *Dev methods have enqueued on this context.
func (g *GPUContext) Sync() error { return gpuErr(g.ctx, C.pgpu_ctx_sync(g.ctx)) }

// The *Dev methods enqueue on the context's stream and return at once; count is in records.

// EncryptWithRDev: c[i] = EncryptWithR(m[i], r[i]) (paillier.go:185-187), n-width m and r, n2-width c.
func (g *GPUContext) EncryptWithRDev(count int, m, r, c *DeviceBuffer) error {
	return gpuErr(g.ctx, C.pgpu_encrypt_with_r_dev(g.ctx, C.size_t(count), m.ptr(), r.ptr(), c.ptr()))
}

// ConstMultDev: out[i] = ConstMult(c[i], k[i]) (operations.go:58-64), k = kBytes-wide little-endian scalars.
func (g *GPUContext) ConstMultDev(count int, c, k *DeviceBuffer, kBytes int, out *DeviceBuffer) error {
	return gpuErr(g.ctx, C.pgpu_const_mult_dev(g.ctx, C.size_t(count), c.ptr(), k.ptr(), C.size_t(kBytes), out.ptr()))
}

// AddPairsDev: out[i] = Add(a[i], b[i]) (operations.go:11-29); out may be a or b.
func (g *GPUContext) AddPairsDev(count int, a, b, out *DeviceBuffer) error {
	return gpuErr(g.ctx, C.pgpu_add_pairs_dev(g.ctx, C.size_t(count), a.ptr(), b.ptr(), out.ptr()))
}

// AddReduceDev: out = Add(c[0], ..., c[count-1]) as one tree reduction.
func (g *GPUContext) AddReduceDev(count int, c, out *DeviceBuffer) error {
	return gpuErr(g.ctx, C.pgpu_add_reduce_dev(g.ctx, C.size_t(count), c.ptr(), out.ptr()))
}

// DecryptDev: m[i] = Decrypt(c[i]) (paillier.go:292-303, CRT over p^2, q^2).
func (g *GPUContext) DecryptDev(count int, c, m *DeviceBuffer) error {
	return gpuErr(g.ctx, C.pgpu_decrypt_dev(g.ctx, C.size_t(count), c.ptr(), m.ptr()))
}

// PartialDecryptDev: out[i] = PartialDecrypt(c[i]) (thresholdkey.go:192-201).
func (g *GPUContext) PartialDecryptDev(count int, c, out *DeviceBuffer) error {
	return gpuErr(g.ctx, C.pgpu_partial_decrypt_dev(g.ctx, C.size_t(count), c.ptr(), out.ptr()))
}

// PinnedBytes is page-locked host memory (pgpu_host_alloc): the host-buffer batch calls copy from / to it at the full
// PCIe rate and overlap the copies of one chunk with the kernels of the next.  The slice aliases C memory: Free it.
type PinnedBytes struct {
	Bytes []byte
	p     unsafe.Pointer
}

// NewPinnedBytes allocates n page-locked bytes.
func NewPinnedBytes(n int) (*PinnedBytes, error) {
	var p unsafe.Pointer
	if rc := C.pgpu_host_alloc(C.size_t(n), &p); rc != C.PGPU_OK {
		return nil, fmt.Errorf("paillier_b200: error %d: %s", int(rc), C.GoString(C.pgpu_last_error(nil)))
	}
	return &PinnedBytes{Bytes: unsafe.Slice((*byte)(p), n), p: p}, nil
}

// Free releases the memory; Bytes must not be used afterwards.
func (b *PinnedBytes) Free() {
	if b.p != nil {
		C.pgpu_host_free(b.p)
		b.p, b.Bytes = nil, nil
	}
}

// CombinePartialDecryptionsZKPBatch = N x ThresholdPublicKey.CombinePartialDecryptionsZKP (thresholdkey.go:164-172):
// shares[j] is server j's batch of proofs, all in the same ciphertext order.  As in the reference the proofs filter per
// ciphertext: ciphertext i is combined from the servers whose proof for i verifies.  errs[i] is the reference's
// "Threshold not meet" where fewer than Threshold remain (plain[i] is nil there).
func (g *GPUContext) CombinePartialDecryptionsZKPBatch(shares [][]*PartialDecryptionZKP) (plain []*gmp.Int, errs []error, err error) {
	k := len(shares)
	if k == 0 {
		return nil, nil, errors.New("Threshold not meet")
	}
	count := len(shares[0])
	ids := make([]C.int, k)
	decs := make([]*gmp.Int, 0, k*count)
	ok := make([]byte, 0, k*count)
	for j, s := range shares {
		if len(s) != count {
			return nil, nil, errors.New("one proof per ciphertext and server")
		}
		if count > 0 {
			ids[j] = C.int(s[0].ID)
		}
		verdicts, verr := g.VerifyProofBatch(s)
		if verr != nil {
			return nil, nil, verr
		}
		for i, p := range s {
			decs = append(decs, p.Decryption)
			if verdicts[i] {
				ok = append(ok, 1)
			} else {
				ok = append(ok, 0)
			}
		}
	}
	d := toRecords(decs, g.wN2)
	m := make([]byte, count*g.wN)
	itemOK := make([]byte, count)
	rc := C.pgpu_combine_verified(g.ctx, C.size_t(count), C.int(k), &ids[0], ptr(d), (*C.uint8_t)(ptr(ok)), ptr(m), (*C.uint8_t)(ptr(itemOK)))
	if rc != C.PGPU_OK && rc != C.PGPU_ERR_THRESHOLD {
		return nil, nil, gpuErr(g.ctx, rc)
	}
	vals := fromRecords(m, g.wN)
	plain, errs = make([]*gmp.Int, count), make([]error, count)
	for i := range vals {
		if itemOK[i] != 0 {
			plain[i] = vals[i]
		} else {
			errs[i] = errors.New("Threshold not meet")
		}
	}
	return plain, errs, nil
}

// ThresholdGroup runs threshold decryption with one share-holder per GPU of this process (pgpu_multi_*, BASELINE
// config 4): PartialDecrypt (+ proofs) of every ciphertext on every device, one NCCL all-gather, proof verification and
// Combine of a ciphertext slice per device.
type ThresholdGroup struct {
	m    *C.pgpu_multi
	ctxs []*GPUContext
}

// NewThresholdGroup takes one context per device, each made from the ThresholdSecretKey of one share-holder of the same key.
func NewThresholdGroup(ctxs []*GPUContext) (*ThresholdGroup, error) {
	raw := make([]*C.pgpu_ctx, len(ctxs))
	for i, g := range ctxs {
		raw[i] = g.ctx
	}
	t := &ThresholdGroup{ctxs: ctxs}
	if rc := C.pgpu_multi_create(&t.m, &raw[0], C.int(len(raw))); rc != C.PGPU_OK {
		return nil, fmt.Errorf("paillier_b200: error %d: %s", int(rc), C.GoString(C.pgpu_multi_last_error(nil)))
	}
	return t, nil
}

// Close destroys the communicators (the contexts stay with their owners).
func (t *ThresholdGroup) Close() {
	if t.m != nil {
		C.pgpu_multi_destroy(t.m)
		t.m = nil
	}
}

// Decrypt recovers the plaintexts of cs.  zkpR is nil (PartialDecrypt + CombinePartialDecryptions) or one slice of
// randomness r in [0, n^2) per share-holder and ciphertext (PartialDecryptionWithZKP + CombinePartialDecryptionsZKP;
// the reference draws r at thresholdkey.go:233).  errs[i] is "Threshold not meet" where too few proofs verified.
func (t *ThresholdGroup) Decrypt(cs []*gmp.Int, zkpR [][]*gmp.Int) (plain []*gmp.Int, errs []error, err error) {
	g := t.ctxs[0]
	count := len(cs)
	c := toRecords(cs, g.wN2)
	m := make([]byte, count*g.wN)
	itemOK := make([]byte, count)
	var rp *unsafe.Pointer
	if zkpR != nil {
		if len(zkpR) != int(C.pgpu_multi_size(t.m)) {
			return nil, nil, errors.New("one slice of randomness per share-holder")
		}
		// C memory for the pointer array: cgo forbids passing a Go slice of Go pointers
		arr := (*[1 << 20]unsafe.Pointer)(C.malloc(C.size_t(len(zkpR)) * C.size_t(unsafe.Sizeof(uintptr(0)))))
		defer C.free(unsafe.Pointer(arr))
		for j, rs := range zkpR {
			rec := C.CBytes(toRecords(rs, g.wN2))
			defer C.free(rec)
			arr[j] = rec
		}
		rp = &arr[0]
	}
	rc := C.pgpu_multi_threshold_round(t.m, C.size_t(count), ptr(c), rp, ptr(m), (*C.uint8_t)(ptr(itemOK)))
	if rc != C.PGPU_OK && rc != C.PGPU_ERR_THRESHOLD {
		return nil, nil, fmt.Errorf("paillier_b200: error %d: %s", int(rc), C.GoString(C.pgpu_multi_last_error(t.m)))
	}
	vals := fromRecords(m, g.wN)
	plain, errs = make([]*gmp.Int, count), make([]error, count)
	for i := range vals {
		if itemOK[i] != 0 {
			plain[i] = vals[i]
		} else {
			errs[i] = errors.New("Threshold not meet")
		}
	}
	return plain, errs, nil
}

// PhasesMs is the device time of the last Decrypt per phase (max over the devices):
// PartialDecrypt, proofs, all-gather, VerifyProof, Combine.
func (t *ThresholdGroup) PhasesMs() [5]float32 {
	var out [5]C.float
	C.pgpu_multi_last_phases_ms(t.m, &out[0])
	return [5]float32{float32(out[0]), float32(out[1]), float32(out[2]), float32(out[3]), float32(out[4])}
}
