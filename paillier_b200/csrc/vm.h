// Micro-program format of the batched modular-exponentiation engine.
//
// Every hot function of the reference is a short sequence of modular
// multiplications and exponentiations whose shape (modulus, exponent or
// window schedule, operand order) is fixed per key and per call
// (/root/reference/paillier.go:206-218, :292-303, thresholdkey.go:192-201 ...).
// The host compiles that shape once into a list of 32-bit ops; the kernel
// (powm_vm in powm.cu) runs the same program for every item of the batch with
// one Montgomery multiplier in the instruction stream, so there is no
// per-item divergence.
#pragma once
#include <cstdint>

namespace pgpu {

enum VmOp : uint32_t {
    OP_END  = 0,
    OP_LDI  = 1,   // x = in[arg][item]                       (plain limbs)
    OP_LDC  = 2,   // x = kconst[arg]                         (per-modulus constant)
    OP_LDT  = 3,   // x = T[arg]                              (per-group scratch table)
    OP_STT  = 4,   // T[arg] = x
    OP_STO  = 5,   // out[arg][item] = x
    OP_SQR  = 6,   // x = mont(x, x), arg times
    OP_MULT = 7,   // x = mont(x, T[arg])
    OP_MULC = 8,   // x = mont(x, kconst[arg])
    OP_MULI = 9,   // x = mont(x, in[arg][item])
    OP_ADDT = 10,  // x = (x + T[arg]) mod n
    OP_ADDC = 11,  // x = (x + kconst[arg]) mod n
    OP_WIN  = 12,  // x = x^(2^w) * T[tbase + bits(exp[item], pos, w)], arg = pos | w<<20 | tbase<<24 (tbase < 8)
    OP_SQMT = 13,  // x = x^(2^nsq) * T[idx], arg = nsq | idx<<12  (sliding-window step)
    OP_FIXW = 14,  // x = x * F[(pos/w)*2^w + bits(exp[item], pos, w)], arg = pos | w<<20  (fixed-base comb step, no squarings)
    OP_SUBT = 15,  // x = (x - T[arg]) mod n
    // chunked access (Montgomery's batch inversion walks a chunk of records per item): record = item*stride + off*S
    OP_LDIO  = 16, // x = in[arg & 3][item, off = arg >> 2]
    OP_MULIO = 17, // x = mont(x, in[arg & 3][item, off = arg >> 2])
    OP_STOO  = 18, // out[arg & 1][item, off = arg >> 2] = x
    // bucket accumulation (Pippenger multi-exponentiation of the encrypted dot product): the group's table is state
    // that persists from item to item
    OP_BKT   = 19, // d = bits(exp[item] + sub*exp_sub, pos, w); T[sub*(2^w+1) + d] = x = mont(x, T[...]), arg = pos | w<<20 | sub<<24
                   // (sub = 0: the dot product's buckets; sub < 8: one bucket set per exponent of a shared-base multi-exponentiation)
};

// 5-bit opcode, 27-bit argument
constexpr uint32_t vm_op(uint32_t code, uint32_t arg) { return (code << 27) | (arg & 0x07ffffffu); }

constexpr int VM_MAX_IN = 4;
constexpr int VM_MAX_OUT = 2;

// Addresses are in 32-bit limbs.  Item i reads in[k] + (i/in_div[k])*in_stride[k] and
// writes out[k] + i*out_stride[k].
struct VmParams {
    const uint32_t* prog;
    uint32_t n_items;
    const uint32_t* mod;               // S limbs
    uint32_t np0;                      // -mod^-1 mod 2^32
    const uint32_t* kconst;            // per-modulus constants, records of S limbs
    const uint32_t* in[VM_MAX_IN];
    uint32_t in_stride[VM_MAX_IN];
    uint32_t in_limbs[VM_MAX_IN];      // limbs actually present per record (<= S, rest reads as 0)
    uint32_t in_div[VM_MAX_IN];        // item i reads record i / in_div (a statement value shared by its proof instances)
    uint32_t* out[VM_MAX_OUT];
    uint32_t out_stride[VM_MAX_OUT];
    uint32_t out_limbs[VM_MAX_OUT];    // limbs stored per record (<= S)
    const uint32_t* exp;               // per-item exponents for OP_WIN / OP_FIXW
    uint32_t exp_stride;               // limbs between exponent records
    uint32_t exp_bits;                 // bits of an exponent record that count (higher bits read as 0)
    uint32_t exp_sub;                  // OP_BKT: limbs between the `sub` exponents of one item
    const uint32_t* fixed;             // fixed-base table for OP_FIXW: records of S limbs, Montgomery form
    uint32_t* table;                   // scratch: [entry][group][S]
    uint32_t n_groups;                 // number of resident groups (table slots)
    uint32_t flags;                    // bit 0: use the general multiplier for squarings too
    uint32_t* dump;                    // n_groups records of S limbs: where idle groups of the last round store
};

}  // namespace pgpu
