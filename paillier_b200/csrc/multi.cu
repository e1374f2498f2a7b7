// pgpu_multi_*: BASELINE config 4 inside the library (SURVEY.md 8b "Threading", 8e): one context per device of ONE
// process, one share-holder per device, NCCL over NVLink for the single exchange of the path.
//
//   device g:  c -> c_g = PartialDecrypt(c) for ALL ciphertexts (thresholdkey.go:192-201) [+ proof, :225-255]
//   ncclAllGather:  [share][ciphertext] on every device (partial decryptions, E, Z)
//   device g:  VerifyProof of every share for ciphertext slice g (:278-311), CombinePartialDecryptionsZKP of the slice
//              with the reference's per-ciphertext filter (:164-172) -> plaintext slice g
//
// A Go caller reaches this with one cgo call (cgo pins the OS thread; the library runs one host thread per device
// underneath because several engine steps synchronise their own stream).  NCCL is bound at run time (dlopen of
// libnccl.so.2): a process that never calls pgpu_multi_create needs no NCCL, and a process that already carries one (the
// Python tests import torch, which bundles its own) keeps using that copy instead of mapping a second one.
#include <dlfcn.h>
#include <nccl.h>

#include <array>
#include <atomic>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <thread>

#include "engine.hpp"

using namespace pgpu;

namespace {

struct NcclApi {
    void* handle = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    std::string why;
    bool ok() const { return handle && CommInitAll && CommDestroy && AllGather && GetErrorString; }
};

NcclApi& nccl() {
    static NcclApi api = [] {
        NcclApi a;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            a.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (a.handle) break;
        }
        if (!a.handle) { a.why = std::string("dlopen(libnccl.so.2): ") + (dlerror() ? dlerror() : "not found"); return a; }
        a.CommInitAll = (decltype(a.CommInitAll))dlsym(a.handle, "ncclCommInitAll");
        a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.handle, "ncclCommDestroy");
        a.AllGather = (decltype(a.AllGather))dlsym(a.handle, "ncclAllGather");
        a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.handle, "ncclGetErrorString");
        a.GetVersion = (decltype(a.GetVersion))dlsym(a.handle, "ncclGetVersion");
        if (!a.ok()) a.why = "libnccl.so.2 lacks ncclCommInitAll / ncclAllGather";
        return a;
    }();
    return api;
}

// all threads of a round meet here; the round goes on only if nobody failed before
struct Rendezvous {
    std::mutex mu;
    std::condition_variable cv;
    int n = 0, arrived = 0, generation = 0;
    std::atomic<int> failed{0};
    bool wait() {
        std::unique_lock<std::mutex> lk(mu);
        const int gen = generation;
        if (++arrived == n) { arrived = 0; ++generation; cv.notify_all(); }
        else cv.wait(lk, [&] { return generation != gen; });
        return failed.load() == 0;
    }
};

// One device's part of a threshold round (runs on its own host thread).
struct DeviceRound {
    pgpu_ctx* ctx;
    int g, G;
    size_t count, lo, n;
    bool zkp;
    uint32_t S = 0, ZL = 0;
    size_t h = 0;
    cudaEvent_t ev[6] = {};
    std::unique_ptr<DevBuf> dc, dec, dr, de, dz, vc, vdec, ve, vz, dok, dm, dit, didx;

    DeviceRound(pgpu_ctx* c, int g_, int G_, size_t count_, bool zkp_) : ctx(c), g(g_), G(G_), count(count_), zkp(zkp_) {
        lo = (count / G) * g + std::min<size_t>(g, count % G);
        n = count / G + ((size_t)g < count % G ? 1 : 0);
    }

    int buf(std::unique_ptr<DevBuf>& b, size_t limbs) {
        b.reset(new DevBuf(ctx, limbs));
        if (b->err != cudaSuccess) return fail(ctx, PGPU_ERR_CUDA, std::string("cudaMallocAsync: ") + cudaGetErrorString(b->err));
        return PGPU_OK;
    }

    int prepare() {
        int rc;
        if ((rc = set_device(ctx))) return rc;
        S = ctx->m_n2.sh.S; ZL = z_limbs(ctx); h = ctx->wn;
        for (auto& e : ev) CU(ctx, cudaEventCreate(&e));
        if ((rc = buf(dc, count * S)) || (rc = buf(dec, (size_t)G * count * S))) return rc;
        if ((rc = buf(dm, std::max<size_t>(n, 1) * h)) || (rc = buf(dit, n / 4 + 2)) || (rc = buf(dok, (size_t)G * n / 4 + 2))) return rc;
        if (zkp) {
            if ((rc = buf(dr, count * S)) || (rc = buf(de, (size_t)G * count * 8)) || (rc = buf(dz, (size_t)G * count * ZL))) return rc;
            if ((rc = buf(vc, (size_t)G * n * S + 1)) || (rc = buf(vdec, (size_t)G * n * S + 1)) || (rc = buf(ve, (size_t)G * n * 8 + 1)) ||
                (rc = buf(vz, (size_t)G * n * ZL + 1))) return rc;
        }
        return PGPU_OK;
    }

    // this share-holder's partial decryptions (and proofs) of EVERY ciphertext, written into its row of the gather buffers
    int before_exchange(const void* c, const void* r) {
        cudaStream_t st = ctx->stream;
        int rc;
        CU(ctx, cudaMemcpyAsync(dc->p, c, count * S * 4, cudaMemcpyHostToDevice, st));
        if (zkp) CU(ctx, cudaMemcpyAsync(dr->p, r, count * S * 4, cudaMemcpyHostToDevice, st));
        CU(ctx, cudaEventRecord(ev[0], st));
        // with proofs the partial decryption comes out of the proof's launch (shared squarings): the "pdec" phase is empty
        if (!zkp && (rc = pdec_dev(ctx, count, dc->p, dec->p + (size_t)g * count * S))) return rc;
        CU(ctx, cudaEventRecord(ev[1], st));
        if (zkp && (rc = zkp_prove_dev(ctx, count, dc->p, dr->p, dec->p + (size_t)g * count * S, de->p + (size_t)g * count * 8,
                                        dz->p + (size_t)g * count * ZL, false))) return rc;
        CU(ctx, cudaEventRecord(ev[2], st));
        return PGPU_OK;
    }

    // the one exchange of the path: in place, every device contributes its row
    int exchange(ncclComm_t comm) {
        NcclApi& N = nccl();
        cudaStream_t st = ctx->stream;
        ncclResult_t r = N.AllGather(dec->p + (size_t)g * count * S, dec->p, count * S * 4, ncclUint8, comm, st);
        if (r == ncclSuccess && zkp) r = N.AllGather(de->p + (size_t)g * count * 8, de->p, count * 32, ncclUint8, comm, st);
        if (r == ncclSuccess && zkp) r = N.AllGather(dz->p + (size_t)g * count * ZL, dz->p, count * ZL * 4, ncclUint8, comm, st);
        if (r != ncclSuccess) return fail(ctx, PGPU_ERR_NCCL, std::string("ncclAllGather: ") + N.GetErrorString(r));
        CU(ctx, cudaEventRecord(ev[3], st));
        return PGPU_OK;
    }

    // VerifyProof of every share for this device's ciphertext slice, Combine of the slice out of the gathered buffer
    int after_exchange(const std::vector<int>& ids, void* plain, uint8_t* item_ok, size_t* n_failed) {
        cudaStream_t st = ctx->stream;
        int rc;
        if (n == 0) { CU(ctx, cudaEventRecord(ev[4], st)); CU(ctx, cudaEventRecord(ev[5], st)); return PGPU_OK; }
        const bool shared = G <= 8 && !getenv("PGPU_NO_SHARED_VERIFY");
        if (zkp && shared) {
            // item-major copies of this slice's proofs (record i*G + s <- gathered row s, ciphertext lo + i): the G proofs of one
            // ciphertext raise the same c^4 to G different Z and share the squarings (zkp_verify_shared_dev)
            std::vector<uint32_t> idx(n * (size_t)G);
            for (size_t i = 0; i < n; ++i) for (int s = 0; s < G; ++s) idx[i * G + s] = (uint32_t)((size_t)s * count + lo + i);
            if ((rc = buf(didx, idx.size()))) return rc;
            if ((rc = upload(ctx, didx->p, idx))) return rc;
            CU(ctx, gather_launch(dec->p, S, didx->p, 1, vdec->p, (uint32_t)idx.size(), st));
            CU(ctx, gather_launch(de->p, 8, didx->p, 1, ve->p, (uint32_t)idx.size(), st));
            CU(ctx, gather_launch(dz->p, ZL, didx->p, 1, vz->p, (uint32_t)idx.size(), st));
            if ((rc = zkp_verify_shared_dev(ctx, n, G, ids.data(), dc->p + lo * S, vdec->p, ve->p, vz->p, (uint8_t*)dok->p))) return rc;
        } else if (zkp) {
            for (int s = 0; s < G; ++s) {
                CU(ctx, cudaMemcpyAsync(vc->p + (size_t)s * n * S, dc->p + lo * S, n * S * 4, cudaMemcpyDeviceToDevice, st));
                CU(ctx, cudaMemcpyAsync(vdec->p + (size_t)s * n * S, dec->p + ((size_t)s * count + lo) * S, n * S * 4, cudaMemcpyDeviceToDevice, st));
                CU(ctx, cudaMemcpyAsync(ve->p + (size_t)s * n * 8, de->p + ((size_t)s * count + lo) * 8, n * 32, cudaMemcpyDeviceToDevice, st));
                CU(ctx, cudaMemcpyAsync(vz->p + (size_t)s * n * ZL, dz->p + ((size_t)s * count + lo) * ZL, n * ZL * 4, cudaMemcpyDeviceToDevice, st));
            }
            if ((rc = zkp_verify_multi_dev(ctx, n, G, ids.data(), vc->p, vdec->p, ve->p, vz->p, (uint8_t*)dok->p))) return rc;
        }
        CU(ctx, cudaEventRecord(ev[4], st));
        if (zkp) rc = combine_verified_dev(ctx, n, G, ids.data(), dec->p + lo * S, count, (const uint8_t*)dok->p, dm->p, (uint8_t*)dit->p, n_failed, shared);
        else rc = combine_dev(ctx, n, G, ids.data(), dec->p + lo * S, dm->p, count);
        if (rc) return rc;
        CU(ctx, cudaEventRecord(ev[5], st));
        CU(ctx, cudaMemcpyAsync((uint8_t*)plain + lo * h * 4, dm->p, n * h * 4, cudaMemcpyDeviceToHost, st));
        if (item_ok) {
            if (zkp) CU(ctx, cudaMemcpyAsync(item_ok + lo, dit->p, n, cudaMemcpyDeviceToHost, st));
            else memset(item_ok + lo, 1, n);
        }
        return PGPU_OK;
    }

    int finish(std::array<float, 5>& ph) {
        const cudaError_t e = cudaStreamSynchronize(ctx->stream);
        int rc = PGPU_OK;
        if (e != cudaSuccess) rc = fail(ctx, PGPU_ERR_CUDA, std::string("threshold round: ") + cudaGetErrorString(e));
        for (int i = 0; i < 5; ++i) {
            float t = 0;
            if (rc == PGPU_OK && ev[i] && ev[i + 1] && cudaEventElapsedTime(&t, ev[i], ev[i + 1]) != cudaSuccess) { t = 0; cudaGetLastError(); }
            ph[i] = t;
        }
        for (auto& x : ev) if (x) cudaEventDestroy(x);
        return rc;
    }
};

}  // namespace

struct pgpu_multi {
    std::vector<pgpu_ctx*> ctx;
    std::vector<int> devices;
    std::vector<ncclComm_t> comms;
    std::string err;
    float phases_ms[5] = {0, 0, 0, 0, 0};      // pdec, prove, all_gather, verify, combine: max over the devices, last round
};

#pragma GCC visibility push(default)
extern "C" {

int pgpu_multi_create(pgpu_multi** out, pgpu_ctx* const* ctxs, int n) {
    if (!out || !ctxs || n < 1) return fail(nullptr, PGPU_ERR_ARG, "pgpu_multi_create: null argument");
    *out = nullptr;
    try {
        for (int i = 0; i < n; ++i) {
            if (!ctxs[i] || !ctxs[i]->has_threshold) return fail(nullptr, PGPU_ERR_STATE, "pgpu_multi_create: every context needs a threshold key");
            if (!(ctxs[i]->n == ctxs[0]->n) || ctxs[i]->tk_l != ctxs[0]->tk_l || ctxs[i]->tk_w != ctxs[0]->tk_w)
                return fail(nullptr, PGPU_ERR_ARG, "pgpu_multi_create: the contexts hold different keys");
            for (int j = 0; j < i; ++j) {
                if (ctxs[j]->device == ctxs[i]->device) return fail(nullptr, PGPU_ERR_ARG, "pgpu_multi_create: one context per device");
                if (ctxs[j]->tk_id == ctxs[i]->tk_id && ctxs[i]->has_share) return fail(nullptr, PGPU_ERR_ARG, "pgpu_multi_create: two contexts hold the same share");
            }
        }
        NcclApi& N = nccl();
        if (!N.ok()) return fail(nullptr, PGPU_ERR_NCCL, "NCCL is not available: " + N.why);
        pgpu_multi* m = new pgpu_multi();
        for (int i = 0; i < n; ++i) { m->ctx.push_back(ctxs[i]); m->devices.push_back(ctxs[i]->device); }
        m->comms.resize(n);
        const ncclResult_t r = N.CommInitAll(m->comms.data(), n, m->devices.data());
        if (r != ncclSuccess) { const std::string msg = std::string("ncclCommInitAll: ") + N.GetErrorString(r); delete m; return fail(nullptr, PGPU_ERR_NCCL, msg); }
        *out = m;
        return PGPU_OK;
    } catch (const std::exception& ex) { return fail(nullptr, PGPU_ERR_ARG, ex.what()); }
}

int pgpu_multi_destroy(pgpu_multi* m) {
    if (!m) return PGPU_OK;
    NcclApi& N = nccl();
    for (size_t i = 0; i < m->comms.size(); ++i) {
        cudaSetDevice(m->devices[i]);
        cudaStreamSynchronize(m->ctx[i]->stream);
        if (N.ok() && m->comms[i]) N.CommDestroy(m->comms[i]);
    }
    delete m;
    return PGPU_OK;
}

const char* pgpu_multi_last_error(const pgpu_multi* m) { return m ? m->err.c_str() : thread_error().c_str(); }

int pgpu_multi_size(const pgpu_multi* m) { return m ? (int)m->ctx.size() : 0; }

int pgpu_multi_last_phases_ms(const pgpu_multi* m, float* out5) {
    if (!m || !out5) return fail(nullptr, PGPU_ERR_ARG, "pgpu_multi_last_phases_ms: null argument");
    for (int i = 0; i < 5; ++i) out5[i] = m->phases_ms[i];
    return PGPU_OK;
}

int pgpu_multi_threshold_round(pgpu_multi* m, size_t count, const void* c, const void* const* zkp_r, void* plain, uint8_t* item_ok) {
    if (!m) return fail(nullptr, PGPU_ERR_ARG, "pgpu_multi_threshold_round: null handle");
    if (count == 0) return PGPU_OK;
    if (!c || !plain) { m->err = "pgpu_multi_threshold_round: null argument"; return fail(nullptr, PGPU_ERR_ARG, m->err); }
    const int G = (int)m->ctx.size();
    for (int g = 0; g < G; ++g) {
        if (!m->ctx[g]->has_share) { m->err = "pgpu_multi_threshold_round: context without a share"; return fail(nullptr, PGPU_ERR_STATE, m->err); }
        if (zkp_r && !zkp_r[g]) { m->err = "pgpu_multi_threshold_round: null randomness for a share-holder"; return fail(nullptr, PGPU_ERR_ARG, m->err); }
    }
    std::vector<int> ids(G);
    for (int g = 0; g < G; ++g) ids[g] = m->ctx[g]->tk_id;
    std::vector<int> rcs(G, PGPU_OK);
    std::vector<std::string> errs(G);
    std::vector<size_t> failed(G, 0);
    std::vector<std::array<float, 5>> phases(G);
    Rendezvous rv; rv.n = G;

    auto worker = [&](int g) {
        DeviceRound R(m->ctx[g], g, G, count, zkp_r != nullptr);
        int rc = R.prepare();
        if (!rc) rc = R.before_exchange(c, zkp_r ? zkp_r[g] : nullptr);
        if (rc) rv.failed++;
        if (rv.wait()) {
            rc = R.exchange(m->comms[g]);
            if (!rc) rc = R.after_exchange(ids, plain, item_ok, &failed[g]);
        } else if (!rc) {
            rc = fail(R.ctx, PGPU_ERR_STATE, "another share-holder failed before the exchange");
        }
        const int rc2 = R.finish(phases[g]);
        if (!rc) rc = rc2;
        rcs[g] = rc;
        if (rc) errs[g] = R.ctx->err;
    };

    std::vector<std::thread> threads;
    for (int g = 0; g < G; ++g) threads.emplace_back(worker, g);
    for (auto& t : threads) t.join();
    size_t total_failed = 0;
    for (int i = 0; i < 5; ++i) m->phases_ms[i] = 0;
    for (int g = 0; g < G; ++g) {
        if (rcs[g]) { m->err = "device " + std::to_string(m->devices[g]) + ": " + errs[g]; return fail(nullptr, rcs[g], m->err); }
        total_failed += failed[g];
        for (int i = 0; i < 5; ++i) m->phases_ms[i] = std::max(m->phases_ms[i], phases[g][i]);
    }
    if (total_failed) {
        m->err = "Threshold not meet for " + std::to_string(total_failed) + " of " + std::to_string(count) + " ciphertexts";
        return fail(nullptr, PGPU_ERR_THRESHOLD, m->err);
    }
    return PGPU_OK;
}

}  // extern "C"
#pragma GCC visibility pop
