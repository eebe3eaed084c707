// spz_comm.cu -- multi-GPU exchange step of the frame path: one process per GPU, NCCL over NVLink / NVSwitch.
//
// The reference is single-threaded host code with no communication of any kind; its path shards at chunk
// granularity (every chunk is an independent reference stream), so encode and decode need no collective.  The one real
// exchange step is gathering the ranks' compressed frames into ONE decodable frame on a destination rank:
//
//   sizes     ncclAllGather of a few u64 per rank (payload offsets at stripe boundaries) on a control communicator
//   placement exclusive scan over the global stripe order (host, a few dozen numbers)
//   payload   ncclGroupStart / ncclSend / ncclRecv / ncclGroupEnd on a data communicator: every stripe's chunk
//             streams, its chunk_off slice and its chunk_nbytes slice land at their final place in the destination
//             frame (no staging copy); the destination's own stripes move with device-to-device copies
//   directory one small kernel rebases the received chunk_off slices and writes the frame header
//
// Streamed form (spz_frame_encode_gather_dev): the shard is encoded stripe by stripe (the slab pipeline of
// spz_api.cu); stripe k's sizes are exchanged and its bytes sent while stripe k+1 is being encoded, so the NVLink
// transfer overlaps the instruction-bound codec kernels.  Stripes are dealt round-robin over the ranks, which is
// what makes every stripe's final position known as soon as its round is encoded.
//
// NCCL is resolved at run time (dlopen "libnccl.so.2": the copy PyTorch already loaded, or the system one), so the
// library has no link-time dependency on it and loads on machines without NCCL.
#include <dlfcn.h>
#include <nccl.h>
#include <stdio.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "spz_host.h"

using namespace spz;

namespace {

struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
    bool load()
    {
        static std::mutex mu;
        std::lock_guard<std::mutex> lk(mu);
        if (ok) return true;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (h) break;
        }
        if (!h) return false;
#define SPZ_SYM(field, sym) *(void**)(&field) = dlsym(h, sym); if (!field) return false
        SPZ_SYM(GetUniqueId, "ncclGetUniqueId");
        SPZ_SYM(CommInitRank, "ncclCommInitRank");
        SPZ_SYM(CommDestroy, "ncclCommDestroy");
        SPZ_SYM(AllGather, "ncclAllGather");
        SPZ_SYM(Send, "ncclSend");
        SPZ_SYM(Recv, "ncclRecv");
        SPZ_SYM(GroupStart, "ncclGroupStart");
        SPZ_SYM(GroupEnd, "ncclGroupEnd");
        SPZ_SYM(GetErrorString, "ncclGetErrorString");
#undef SPZ_SYM
        ok = true;
        return true;
    }
} g_nccl;

int nccl_fail(ncclResult_t r, const char* where)
{
    char msg[200];
    snprintf(msg, sizeof(msg), "%s: NCCL: %s", where, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
    return fail_msg(SPZ_E_CUDA, msg);
}
#define NK(call) do { ncclResult_t r__ = (call); if (r__ != ncclSuccess) return nccl_fail(r__, #call); } while (0)

constexpr int kMaxLocalStripes = 62;              // stripes per rank in one gather (control record = 64 u64)
constexpr int kCtlWords = kMaxLocalStripes + 2;   // [0] codec fingerprint, [1..] payload offset at local stripe boundaries
constexpr int kMaxRanks = 64;
constexpr int kMaxSegs = kMaxLocalStripes * kMaxRanks;

struct Comm {
    ncclComm_t data = nullptr, ctl = nullptr;
    int nranks = 0, rank = 0;
    cudaStream_t s_ctl = nullptr, s_data = nullptr;
    cudaEvent_t ev_main = nullptr, ev_data = nullptr;
    cudaEvent_t ev_c[kMaxLocalStripes], ev_sz[kMaxLocalStripes];
    unsigned long long* d_send = nullptr;      // kCtlWords
    unsigned long long* d_recv = nullptr;      // nranks * kCtlWords (non-streamed)  |  kMaxLocalStripes * nranks (streamed, one word per rank and round)
    unsigned long long* h_recv = nullptr;      // pinned mirror of d_recv
    unsigned long long* d_seg = nullptr;       // seg_lo[kMaxSegs + 1] | seg_delta[kMaxSegs]
    unsigned long long* h_seg = nullptr;       // pinned
    std::mutex mu;
    // peer-memory gather (spz_comm_shared_alloc / spz_frame_encode_gather_p2p_dev)
    uint8_t* sh_local = nullptr;               // my shared region: [SharedCtl (kSharedCtl bytes)][gathered frame area]
    uint8_t* sh_peer[64] = {nullptr};          // every rank's region mapped here (own rank: sh_local)
    uint64_t sh_bytes[64] = {0};
    cudaStream_t s_p2p = nullptr;              // high priority: the transfer CTAs slip in between the encoder's
    cudaStream_t s_enc[2] = {nullptr, nullptr};// stripes are encoded on three streams in turn (a stripe is less than a machine round)
    uint32_t epoch = 0;
};

// ------------------------------------------------------------------ partition of the global stream
struct Seg { int rank; uint64_t local_lo, nchunks, global_lo; };
struct Plan {
    FrameLayout f;                         // of the global stream
    std::vector<Seg> segs;                 // global order
    std::vector<uint64_t> local_chunks;    // per rank
};

bool make_plan(Plan& pl, uint64_t total_len, uint16_t ndims, uint32_t chunk_rows, uint32_t stripe_chunks, int nranks)
{
    pl.f = frame_layout(total_len, ndims, chunk_rows);
    pl.segs.clear();
    pl.local_chunks.assign(nranks, 0);
    const uint64_t nch = pl.f.nchunks;
    if (stripe_chunks == 0) {                  // contiguous, balanced ranges of whole chunks
        const uint64_t base = nch / nranks, extra = nch % nranks;
        uint64_t lo = 0;
        for (int r = 0; r < nranks; r++) {
            const uint64_t n = base + ((uint64_t)r < extra ? 1 : 0);
            if (n) pl.segs.push_back(Seg{r, 0, n, lo});
            pl.local_chunks[r] = n;
            lo += n;
        }
    } else {                                   // stripes dealt round-robin
        const uint64_t S = stripe_chunks;
        const uint64_t nstripes = (nch + S - 1) / S;
        for (uint64_t j = 0; j < nstripes; j++) {
            const int r = (int)(j % (uint64_t)nranks);
            const uint64_t n = (j + 1) * S <= nch ? S : nch - j * S;
            pl.segs.push_back(Seg{r, pl.local_chunks[r], n, j * S});
            pl.local_chunks[r] += n;
        }
    }
    if (pl.segs.size() > (size_t)kMaxSegs) return false;
    for (int r = 0; r < nranks; r++) {
        uint64_t cnt = 0;
        for (const Seg& s : pl.segs) if (s.rank == r) cnt++;
        if (cnt > (uint64_t)kMaxLocalStripes) return false;
    }
    return true;
}

uint64_t seg_elems(const Plan& pl, const Seg& s, uint64_t total_len)
{
    const uint64_t a = s.global_lo * pl.f.chunk_len, b = (s.global_lo + s.nchunks) * pl.f.chunk_len;
    return (b < total_len ? b : total_len) - (a < total_len ? a : total_len);
}

uint64_t plan_local_len(const Plan& pl, int rank, uint64_t total_len)
{
    uint64_t n = 0;
    for (const Seg& s : pl.segs) if (s.rank == rank) n += seg_elems(pl, s, total_len);
    return n;
}

__global__ void spz_pick_offsets_kernel(const uint64_t* __restrict__ chunk_off, const uint64_t* __restrict__ idx, int n,
                                        unsigned long long* out)
{
    const int i = threadIdx.x;
    if (i < n) out[i] = chunk_off[idx[i]];
}

uint64_t codec_fingerprint(const spz_frame_header& h)
{
    return ((uint64_t)h.predictor << 56) | ((uint64_t)h.elem_sz << 48) | ((uint64_t)(h.ndims & 0xffffu) << 32) | h.chunk_rows;
}

// Posts the transfers of the segments [s0, s1) of the plan (global order) and records their placement.
//   bnd[r][k] = payload offset inside rank r's local frame at the start of its k-th local segment
//   G         = running payload offset inside the gathered frame (in/out)
// Returns the rebase entries (seg_lo, seg_delta) through lo_out / delta_out at index s.
int post_segments(Comm& c, int dst, const Plan& pl, size_t s0, size_t s1, const std::vector<std::vector<uint64_t>>& bnd,
                  std::vector<size_t>& local_idx, uint64_t& G, const uint8_t* local_frame, const FrameLayout& lf,
                  uint8_t* out_frame, uint64_t out_cap, unsigned long long* lo_out, long long* delta_out, cudaStream_t st)
{
    const FrameLayout& gf = pl.f;
    // this rank joins the NCCL group iff it has something to send (owns a remote-bound segment) or to receive
    bool in_group = false;
    for (size_t s = s0; s < s1; s++) {
        const int owner = pl.segs[s].rank;
        if (owner != dst && (c.rank == dst || owner == c.rank)) in_group = true;
    }
    if (in_group) NK(g_nccl.GroupStart());
    int rc = SPZ_OK;
    for (size_t s = s0; s < s1 && rc == SPZ_OK; s++) {
        const Seg& sg = pl.segs[s];
        const size_t k = local_idx[sg.rank]++;
        const uint64_t L0 = bnd[sg.rank][k], L1 = bnd[sg.rank][k + 1];
        if (L1 < L0 || (L0 & 15ull)) { rc = fail_msg(SPZ_E_STREAM, "gather: inconsistent stripe boundaries"); break; }
        const uint64_t nbytes = L1 - L0;
        lo_out[s] = sg.global_lo;
        delta_out[s] = (long long)G - (long long)L0;
        if (c.rank == dst) {
            if (gf.payload_off + G + nbytes + 16 > out_cap) { rc = SPZ_E_CAPACITY; break; }
            uint8_t* p_pay = out_frame + gf.payload_off + G;
            uint8_t* p_off = out_frame + gf.off_dir + 8 * sg.global_lo;
            uint8_t* p_nb = out_frame + gf.off_sizes + 4 * sg.global_lo;
            if (sg.rank == dst) {
                if (nbytes) CK(cudaMemcpyAsync(p_pay, local_frame + lf.payload_off + L0, nbytes, cudaMemcpyDeviceToDevice, st));
                CK(cudaMemcpyAsync(p_off, local_frame + lf.off_dir + 8 * sg.local_lo, 8 * sg.nchunks, cudaMemcpyDeviceToDevice, st));
                CK(cudaMemcpyAsync(p_nb, local_frame + lf.off_sizes + 4 * sg.local_lo, 4 * sg.nchunks, cudaMemcpyDeviceToDevice, st));
            } else {
                if (nbytes) NK(g_nccl.Recv(p_pay, nbytes, ncclUint8, sg.rank, c.data, st));
                NK(g_nccl.Recv(p_off, 8 * sg.nchunks, ncclUint8, sg.rank, c.data, st));
                NK(g_nccl.Recv(p_nb, 4 * sg.nchunks, ncclUint8, sg.rank, c.data, st));
            }
        } else if (sg.rank == c.rank) {
            if (nbytes) NK(g_nccl.Send(local_frame + lf.payload_off + L0, nbytes, ncclUint8, dst, c.data, st));
            NK(g_nccl.Send(local_frame + lf.off_dir + 8 * sg.local_lo, 8 * sg.nchunks, ncclUint8, dst, c.data, st));
            NK(g_nccl.Send(local_frame + lf.off_sizes + 4 * sg.local_lo, 4 * sg.nchunks, ncclUint8, dst, c.data, st));
        }
        G += nbytes;
    }
    if (in_group) { const ncclResult_t r = g_nccl.GroupEnd(); if (r != ncclSuccess && rc == SPZ_OK) return nccl_fail(r, "ncclGroupEnd"); }
    return rc;
}

// destination: rebase the received chunk_off slices and write the gathered frame's header
int finish_gathered(Comm& c, const Plan& pl, const spz_frame_header& local_h, uint64_t total_len, uint64_t G,
                    const unsigned long long* lo, const long long* delta, uint8_t* out_frame, cudaStream_t st)
{
    const size_t ns = pl.segs.size();
    memcpy(c.h_seg, lo, 8 * ns);
    c.h_seg[ns] = pl.f.nchunks;
    memcpy(c.h_seg + kMaxSegs + 1, delta, 8 * ns);
    CK(cudaMemcpyAsync(c.d_seg, c.h_seg, 8 * (2 * kMaxSegs + 1), cudaMemcpyHostToDevice, st));
    spz_frame_header h = local_h;
    h.len = total_len; h.nchunks = pl.f.nchunks; h.payload_off = pl.f.payload_off;
    CK(launch_rebase_directory((uint64_t*)(out_frame + pl.f.off_dir), pl.f.nchunks, (const uint64_t*)c.d_seg,
                               (const int64_t*)(c.d_seg + kMaxSegs + 1), (uint32_t)ns, G,
                               (uint32_t*)(out_frame + pl.f.off_sizes), (spz_frame_header*)out_frame, h, st));
    return SPZ_OK;
}

// ------------------------------------------------------------------ peer-memory gather: kernels
// Every rank owns a shared region (cudaMalloc + CUDA IPC, mapped by all peers): a control block the peers write their
// stripe sizes and completion flags into, followed (on the destination) by the gathered frame.  Words are tagged with
// the call's epoch in their top 16 bits, so nothing has to be cleared between calls.
constexpr uint64_t kSharedCtl = 64 * 1024;
constexpr long long kSpinCycles = 8000000000ll;         // ~4 s: a wait that long means a peer failed; the frame is flagged invalid
struct SharedCtl {
    unsigned long long tot[kMaxLocalStripes][kMaxRanks];    // [round][rank]: tag << 48 | payload bytes of that rank's stripe of the round
    unsigned long long done[kMaxRanks];                     // destination only: tag << 48 | error bits of rank
    unsigned long long place[kMaxLocalStripes];             // local: offset of my stripe of round k inside the gathered payload
    unsigned int err;                                       // local: 1 wait timed out | 2 destination too small
};
static_assert(sizeof(SharedCtl) <= kSharedCtl, "control block too large");
struct PeerTab { SharedCtl* ctl[kMaxRanks]; };
constexpr unsigned long long kValMask = (1ull << 48) - 1ull;

__device__ __forceinline__ unsigned long long ld_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ bool wait_tag(const unsigned long long* p, unsigned long long tag, unsigned long long& val)
{
    const long long t0 = clock64();
    for (;;) {
        const unsigned long long v = ld_sys(p);
        if ((v >> 48) == tag) { val = v & kValMask; return true; }
        if (clock64() - t0 > kSpinCycles) return false;
        __nanosleep(256);
    }
}

// my stripe of `round` is scanned: tell every rank its payload size (0 if I own no stripe in this round)
__global__ void spz_p2p_publish_kernel(PeerTab pt, int nranks, int rank, int round, const unsigned long long* d_tot, int has,
                                       unsigned long long tag)
{
    const int i = threadIdx.x;
    if (i < nranks) {
        const unsigned long long v = has ? d_tot[round + 1] - d_tot[round] : 0ull;
        st_sys(&pt.ctl[i]->tot[round][rank], (tag << 48) | v);
    }
}

// place[round] = bytes of every stripe that precedes mine in global order (round-major, rank-minor)
__global__ void spz_p2p_place_kernel(SharedCtl* me, int nranks, int rank, int round, unsigned long long tag)
{
    __shared__ unsigned long long sum;
    __shared__ int bad;
    if (threadIdx.x == 0) { sum = 0; bad = 0; }
    __syncthreads();
    const int i = threadIdx.x;
    if (i < nranks) {
        unsigned long long add = 0, v = 0;
        bool ok = true;
        if (round > 0 && i >= rank) { ok = wait_tag(&me->tot[round - 1][i], tag, v); add += v; }
        if (ok && i < rank) { ok = wait_tag(&me->tot[round][i], tag, v); add += v; }
        if (ok) atomicAdd(&sum, add); else bad = 1;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long prev = round > 0 ? me->place[round - 1] : 0ull;
        if (bad || prev == ~0ull) { me->err |= 1u; me->place[round] = ~0ull; }
        else me->place[round] = prev + sum;
    }
}

// Compaction fused with the transfer: chunk c's stream moves from its worst-case slot in LOCAL memory straight to its
// final place in the DESTINATION's gathered frame (peer memory: the stores travel over NVLink), and the chunk's
// directory entry with it.  One warp per chunk, 128-bit copies, four loads in flight per lane; a bounded grid, so the
// transfer only takes a few CTA slots from the encoder that runs beside it.
__global__ void __launch_bounds__(256)
spz_p2p_gather_kernel(const uint8_t* __restrict__ slots, uint64_t slot_stride, uint64_t nchunks,
                      const uint64_t* __restrict__ chunk_off, const uint32_t* __restrict__ chunk_nbytes,
                      const unsigned long long* __restrict__ local_base, SharedCtl* me, int round,
                      uint8_t* __restrict__ dst_payload, uint64_t dst_payload_cap, uint64_t* __restrict__ dst_off,
                      uint32_t* __restrict__ dst_nb)
{
    const unsigned long long place = me->place[round];
    if (place == ~0ull) return;
    const unsigned long long lb = local_base[0], le = local_base[1];
    if (place + (le - lb) + 16ull > dst_payload_cap) {
        if (threadIdx.x == 0 && blockIdx.x == 0) me->err |= 2u;
        return;
    }
    const unsigned long long delta = place - lb;
    const int lane = threadIdx.x & 31;
    const uint64_t wstride = (uint64_t)gridDim.x * (blockDim.x >> 5);
    for (uint64_t c = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); c < nchunks; c += wstride) {
        const uint32_t nb = chunk_nbytes[c];
        const uint32_t units = (nb + 15u) >> 4;
        const unsigned long long off = chunk_off[c] + delta;
        const uint4* __restrict__ s4 = reinterpret_cast<const uint4*>(slots + c * slot_stride);
        uint4* __restrict__ d4 = reinterpret_cast<uint4*>(dst_payload + off);
        if (lane == 0) { dst_off[c] = off; dst_nb[c] = nb; }
        for (uint32_t u0 = 0; u0 < units; u0 += 128u) {
            uint4 x[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t u = u0 + (uint32_t)lane + 32u * k;
                if (u < units) x[k] = s4[u];
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t u = u0 + (uint32_t)lane + 32u * k;
                if (u < units) {
                    uint4 v = x[k];
                    if (u == units - 1 && (nb & 15u)) {     // bytes past the stream's true end are written as zero
                        uint32_t w[4] = {v.x, v.y, v.z, v.w};
                        const uint32_t keep = nb & 15u;
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const uint32_t lo = 4u * q;
                            if (keep <= lo) w[q] = 0;
                            else if (keep < lo + 4) w[q] &= (1u << (8 * (keep - lo))) - 1u;
                        }
                        v = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                    d4[u] = v;
                }
            }
        }
    }
}

// all my stripes have left: tell the destination (after a system-scope fence: the flag must not overtake the data)
__global__ void spz_p2p_signal_kernel(SharedCtl* me, const unsigned int* enc_err, SharedCtl* dst, int rank, unsigned long long tag)
{
    const unsigned long long e = (unsigned long long)me->err | (*enc_err ? 4ull : 0ull);     // 4: a chunk did not fit its slot
    __threadfence_system();
    st_sys(&dst->done[rank], (tag << 48) | e);
}

// destination: wait for every rank, then write the directory terminator, the read slack and the frame header
__global__ void spz_p2p_finish_kernel(SharedCtl* me, int nranks, int nrounds, unsigned long long tag, uint64_t* chunk_off,
                                      uint64_t nchunks, uint32_t* chunk_nbytes, spz_frame_header* hdr, spz_frame_header hdr_in)
{
    __shared__ unsigned long long total;
    __shared__ unsigned int err;
    if (threadIdx.x == 0) { total = 0; err = 0; }
    __syncthreads();
    const int i = threadIdx.x;
    if (i < nranks) {
        unsigned long long v = 0, sum = 0;
        if (!wait_tag(&me->done[i], tag, v)) atomicOr(&err, 1u);
        else {
            if (v) atomicOr(&err, (unsigned int)v);
            for (int k = 0; k < nrounds; k++) sum += ld_sys(&me->tot[k][i]) & kValMask;
            atomicAdd(&total, sum);
        }
    }
    __syncthreads();
    __threadfence_system();
    if (threadIdx.x == 0) {
        chunk_off[nchunks] = total;
        uint8_t* gap = reinterpret_cast<uint8_t*>(chunk_nbytes + nchunks);
        uint8_t* pay = reinterpret_cast<uint8_t*>(hdr) + hdr_in.payload_off;
        while (gap < pay) *gap++ = 0;
        for (int b = 0; b < 16; b++) pay[total + b] = 0;                 // read slack
        spz_frame_header h = hdr_in;
        h.payload_bytes = total;
        h.frame_bytes = h.payload_off + total + 16;
        if (err) h.magic = 0;                                           // a rank failed: the frame is not valid
        *hdr = h;
    }
}

void release_shared(Comm& c)
{
    if (!c.sh_local) return;
    cudaDeviceSynchronize();
    for (int i = 0; i < c.nranks; i++)
        if (i != c.rank && c.sh_peer[i]) cudaIpcCloseMemHandle(c.sh_peer[i]);
    cudaFree(c.sh_local);
    c.sh_local = nullptr;
    for (int i = 0; i < 64; i++) { c.sh_peer[i] = nullptr; c.sh_bytes[i] = 0; }
}

}  // namespace

// ------------------------------------------------------------------ C ABI

extern "C" int spz_comm_unique_id(void* id)
{
    if (!id) return SPZ_E_ARG;
    if (!g_nccl.load()) return fail_msg(SPZ_E_UNSUPPORTED, "libnccl.so.2 not found");
    ncclUniqueId a, b;
    NK(g_nccl.GetUniqueId(&a));
    NK(g_nccl.GetUniqueId(&b));
    static_assert(2 * sizeof(ncclUniqueId) <= SPZ_COMM_ID_BYTES, "id blob too small");
    memcpy(id, &a, sizeof(a));
    memcpy((char*)id + sizeof(a), &b, sizeof(b));
    return SPZ_OK;
}

extern "C" int spz_comm_init(spz_comm_t* comm, int nranks, int rank, const void* id)
{
    if (!comm || !id || nranks < 1 || nranks > kMaxRanks || rank < 0 || rank >= nranks) return SPZ_E_ARG;
    if (!g_nccl.load()) return fail_msg(SPZ_E_UNSUPPORTED, "libnccl.so.2 not found");
    Comm* c = new Comm();
    c->nranks = nranks; c->rank = rank;
    ncclUniqueId a, b;
    memcpy(&a, id, sizeof(a));
    memcpy(&b, (const char*)id + sizeof(a), sizeof(b));
    NK(g_nccl.CommInitRank(&c->data, nranks, a, rank));
    NK(g_nccl.CommInitRank(&c->ctl, nranks, b, rank));
    CK(cudaStreamCreateWithFlags(&c->s_ctl, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->s_data, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c->ev_main, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_data, cudaEventDisableTiming));
    for (int i = 0; i < kMaxLocalStripes; i++) {
        CK(cudaEventCreateWithFlags(&c->ev_c[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c->ev_sz[i], cudaEventDisableTiming));
    }
    const size_t recv_words = (size_t)nranks * kCtlWords;
    CK(cudaMalloc(&c->d_send, 8 * (kCtlWords + kMaxLocalStripes + 2)));       // control record | stripe boundary indices
    CK(cudaMalloc(&c->d_recv, 8 * recv_words));
    CK(cudaMallocHost(&c->h_recv, 8 * recv_words));
    CK(cudaMalloc(&c->d_seg, 8 * (2 * kMaxSegs + 1)));
    CK(cudaMallocHost(&c->h_seg, 8 * (2 * kMaxSegs + 1)));
    *comm = c;
    return SPZ_OK;
}

extern "C" int spz_comm_destroy(spz_comm_t comm)
{
    Comm* c = (Comm*)comm;
    if (!c) return SPZ_OK;
    cudaDeviceSynchronize();
    release_shared(*c);
    if (c->s_p2p) { cudaStreamDestroy(c->s_p2p); cudaStreamDestroy(c->s_enc[0]); cudaStreamDestroy(c->s_enc[1]); }
    if (c->data) g_nccl.CommDestroy(c->data);
    if (c->ctl) g_nccl.CommDestroy(c->ctl);
    if (c->s_ctl) cudaStreamDestroy(c->s_ctl);
    if (c->s_data) cudaStreamDestroy(c->s_data);
    if (c->ev_main) { cudaEventDestroy(c->ev_main); cudaEventDestroy(c->ev_data); }
    for (int i = 0; i < kMaxLocalStripes; i++) { if (c->ev_c[i]) cudaEventDestroy(c->ev_c[i]); if (c->ev_sz[i]) cudaEventDestroy(c->ev_sz[i]); }
    cudaFree(c->d_send); cudaFree(c->d_recv); cudaFreeHost(c->h_recv); cudaFree(c->d_seg); cudaFreeHost(c->h_seg);
    delete c;
    return SPZ_OK;
}

extern "C" uint64_t spz_shard_len(uint64_t total_len, uint16_t ndims, uint32_t chunk_rows, uint32_t stripe_chunks, int nranks,
                                  int rank)
{
    if (ndims == 0 || chunk_rows == 0 || nranks < 1 || rank < 0 || rank >= nranks) return 0;
    Plan pl;
    if (!make_plan(pl, total_len, ndims, chunk_rows, stripe_chunks, nranks)) return 0;
    return plan_local_len(pl, rank, total_len);
}

extern "C" uint64_t spz_frame_gather_max_bytes(int elem_sz, uint64_t total_len, uint16_t ndims, uint32_t chunk_rows)
{
    return spz_frame_max_bytes(elem_sz, total_len, ndims, chunk_rows);
}

extern "C" int spz_frame_gather_dev(spz_comm_t comm, int dst, uint64_t total_len, uint32_t stripe_chunks, const void* d_frame,
                                    void* d_gathered, uint64_t gathered_cap, spz_stream_t stream_)
{
    Comm* c = (Comm*)comm;
    if (!c || dst < 0 || dst >= c->nranks || !d_frame) return SPZ_E_ARG;
    if (c->rank == dst && !d_gathered) return SPZ_E_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    cudaStream_t st = (cudaStream_t)stream_;
    spz_frame_header h;
    CK(cudaMemcpyAsync(&h, d_frame, sizeof(h), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (h.magic != SPZ_FRAME_MAGIC) return fail_msg(SPZ_E_STREAM, "gather: the local frame is not valid");
    Plan pl;
    if (!make_plan(pl, total_len, (uint16_t)h.ndims, h.chunk_rows, stripe_chunks, c->nranks)) return SPZ_E_UNSUPPORTED;
    if (plan_local_len(pl, c->rank, total_len) != h.len || pl.local_chunks[c->rank] != h.nchunks)
        return fail_msg(SPZ_E_ARG, "gather: the local frame does not cover this rank's shard of total_len");
    const FrameLayout lf = frame_layout(h.len, (uint16_t)h.ndims, h.chunk_rows);
    const uint8_t* lframe = (const uint8_t*)d_frame;

    // ---- control record: fingerprint + payload offsets at my stripe boundaries
    std::vector<uint64_t> bidx;
    for (const Seg& s : pl.segs) if (s.rank == c->rank) bidx.push_back(s.local_lo);
    bidx.push_back(h.nchunks);
    const int nb = (int)bidx.size();
    unsigned long long* d_idx = c->d_send + kCtlWords;
    CK(cudaMemsetAsync(c->d_send, 0, 8 * kCtlWords, st));
    const uint64_t fp = codec_fingerprint(h);
    CK(cudaMemcpyAsync(c->d_send, &fp, 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_idx, bidx.data(), 8 * nb, cudaMemcpyHostToDevice, st));
    count_launch();
    spz_pick_offsets_kernel<<<1, 64, 0, st>>>((const uint64_t*)(lframe + lf.off_dir), (const uint64_t*)d_idx, nb, c->d_send + 1);
    CK(cudaGetLastError());
    NK(g_nccl.AllGather(c->d_send, c->d_recv, kCtlWords, ncclUint64, c->ctl, st));
    CK(cudaMemcpyAsync(c->h_recv, c->d_recv, 8 * (size_t)c->nranks * kCtlWords, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    std::vector<std::vector<uint64_t>> bnd(c->nranks);
    for (int r = 0; r < c->nranks; r++) {
        if (c->h_recv[(size_t)r * kCtlWords] != fp) return fail_msg(SPZ_E_ARG, "gather: ranks disagree on the codec parameters");
        size_t cnt = 0;
        for (const Seg& s : pl.segs) if (s.rank == r) cnt++;
        bnd[r].assign(c->h_recv + (size_t)r * kCtlWords + 1, c->h_recv + (size_t)r * kCtlWords + 2 + cnt);
    }
    // ---- payload + directory slices, then the directory rebase on the destination
    std::vector<size_t> lidx(c->nranks, 0);
    std::vector<unsigned long long> lo(pl.segs.size());
    std::vector<long long> delta(pl.segs.size());
    uint64_t G = 0;
    int rc = post_segments(*c, dst, pl, 0, pl.segs.size(), bnd, lidx, G, lframe, lf, (uint8_t*)d_gathered, gathered_cap,
                           lo.data(), delta.data(), st);
    if (rc != SPZ_OK) return rc;
    if (c->rank == dst) rc = finish_gathered(*c, pl, h, total_len, G, lo.data(), delta.data(), (uint8_t*)d_gathered, st);
    return rc;
}

extern "C" int spz_frame_encode_gather_dev(spz_comm_t comm, int dst, int predictor, int elem_sz, const void* d_src,
                                           uint64_t total_len, uint16_t ndims, uint32_t chunk_rows, uint32_t stripe_chunks,
                                           void* d_frame, uint64_t frame_cap, void* d_work, uint64_t work_bytes,
                                           void* d_gathered, uint64_t gathered_cap, spz_stream_t stream_)
{
    Comm* c = (Comm*)comm;
    if (!c || dst < 0 || dst >= c->nranks || stripe_chunks == 0) return SPZ_E_ARG;
    if (c->rank == dst && !d_gathered) return SPZ_E_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    cudaStream_t st = (cudaStream_t)stream_;
    Plan pl;
    if (ndims == 0 || chunk_rows == 0) return SPZ_E_ARG;
    if (!make_plan(pl, total_len, ndims, chunk_rows, stripe_chunks, c->nranks)) return SPZ_E_UNSUPPORTED;
    const uint64_t len = plan_local_len(pl, c->rank, total_len);
    if (len == 0) return fail_msg(SPZ_E_ARG, "encode_gather: every rank must own at least one stripe");
    FrameJob j;
    int rc = frame_job_init(j, predictor, elem_sz, d_src, len, ndims, chunk_rows, d_frame, frame_cap, d_work, work_bytes,
                            nullptr, stripe_chunks);
    if (rc != SPZ_OK) return rc;
    if (j.slab_chunks != stripe_chunks || j.f.nchunks != pl.local_chunks[c->rank]) return SPZ_E_ARG;
    const FrameLayout& lf = j.f;
    uint8_t* lframe = (uint8_t*)d_frame;
    const uint64_t nrounds = (pl.segs.size() + c->nranks - 1) / c->nranks;     // global rounds; I own a stripe in round k iff k < j.nslabs
    if (nrounds > (uint64_t)kMaxLocalStripes) return SPZ_E_UNSUPPORTED;
    if ((rc = frame_job_reset(j, st)) != SPZ_OK) return rc;
    SidePipe& sp = side_pipe();
    CK(sp.init((int)j.nslabs));
    CK(cudaEventRecord(c->ev_main, st));
    CK(cudaStreamWaitEvent(sp.side, c->ev_main, 0));
    CK(cudaStreamWaitEvent(c->s_ctl, c->ev_main, 0));
    CK(cudaStreamWaitEvent(c->s_data, c->ev_main, 0));
    CK(cudaMemsetAsync(c->d_send, 0, 8 * kCtlWords, c->s_ctl));

    spz_frame_header h = j.h;
    std::vector<std::vector<uint64_t>> bnd(c->nranks, std::vector<uint64_t>(1, 0));   // running payload offsets per rank
    std::vector<size_t> lidx(c->nranks, 0);
    std::vector<unsigned long long> lo(pl.segs.size());
    std::vector<long long> delta(pl.segs.size());
    uint64_t G = 0;

    auto enqueue_encode = [&](uint64_t k) -> int {
        // codec kernel on the caller's stream; directory scan + compaction on the side stream (overlaps the next slab)
        if (k < j.nslabs) {
            if (int r = frame_job_encode_slab(j, k, st)) return r;
            CK(cudaEventRecord(sp.ev_enc[k], st));
            CK(cudaStreamWaitEvent(sp.side, sp.ev_enc[k], 0));
            if (int r = frame_job_compact_slab(j, k, sp.side)) return r;
            CK(cudaEventRecord(c->ev_c[k], sp.side));
        }
        return SPZ_OK;
    };
    if ((rc = enqueue_encode(0)) != SPZ_OK) return rc;
    for (uint64_t k = 0; k < nrounds; k++) {
        // sizes of round k: everyone's running payload total after its k-th stripe (ranks without one repeat the last)
        const uint64_t kk = k < j.nslabs ? k : (j.nslabs ? j.nslabs - 1 : 0);
        if (j.nslabs) CK(cudaStreamWaitEvent(c->s_ctl, c->ev_c[kk], 0));
        NK(g_nccl.AllGather(j.d_tot() + kk + 1, c->d_recv + k * c->nranks, 1, ncclUint64, c->ctl, c->s_ctl));
        CK(cudaMemcpyAsync(c->h_recv + k * c->nranks, c->d_recv + k * c->nranks, 8 * (size_t)c->nranks, cudaMemcpyDeviceToHost, c->s_ctl));
        CK(cudaEventRecord(c->ev_sz[k], c->s_ctl));
        if ((rc = enqueue_encode(k + 1)) != SPZ_OK) return rc;          // keeps the GPU busy while the host waits for the sizes
        CK(cudaEventSynchronize(c->ev_sz[k]));
        const size_t s0 = (size_t)k * c->nranks, s1 = (s0 + c->nranks < pl.segs.size()) ? s0 + c->nranks : pl.segs.size();
        for (size_t s = s0; s < s1; s++) bnd[pl.segs[s].rank].push_back(c->h_recv[k * c->nranks + pl.segs[s].rank]);
        if (k < j.nslabs) CK(cudaStreamWaitEvent(c->s_data, c->ev_c[k], 0));       // my stripe k is compacted
        rc = post_segments(*c, dst, pl, s0, s1, bnd, lidx, G, lframe, lf, (uint8_t*)d_gathered, gathered_cap, lo.data(),
                           delta.data(), c->s_data);
        if (rc != SPZ_OK) { cudaDeviceSynchronize(); return rc; }
    }
    CK(cudaEventRecord(c->ev_data, c->s_data));
    CK(cudaStreamWaitEvent(st, c->ev_data, 0));
    if (j.nslabs) CK(cudaStreamWaitEvent(st, c->ev_c[j.nslabs - 1], 0));
    if (c->rank == dst) rc = finish_gathered(*c, pl, h, total_len, G, lo.data(), delta.data(), (uint8_t*)d_gathered, st);
    return rc;
}

// ------------------------------------------------------------------ peer-memory gather: C ABI

extern "C" int spz_comm_shared_alloc(spz_comm_t comm, uint64_t gathered_bytes, void** d_gathered)
{
    Comm* c = (Comm*)comm;
    if (!c) return SPZ_E_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    release_shared(*c);
    if (!c->s_p2p) {
        int lo = 0, hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CK(cudaStreamCreateWithPriority(&c->s_p2p, cudaStreamNonBlocking, hi));
        CK(cudaStreamCreateWithFlags(&c->s_enc[0], cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&c->s_enc[1], cudaStreamNonBlocking));
    }
    const uint64_t bytes = kSharedCtl + ((gathered_bytes + 255ull) & ~255ull);
    CK(cudaMalloc(&c->sh_local, bytes));
    CK(cudaMemset(c->sh_local, 0, kSharedCtl));
    CK(cudaDeviceSynchronize());                                        // the control block is clean before anyone can see it
    // exchange [IPC handle (64 bytes) | region bytes] over the control communicator
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    unsigned long long words[16] = {0};
    cudaIpcMemHandle_t mine;
    CK(cudaIpcGetMemHandle(&mine, c->sh_local));
    memcpy(words, &mine, sizeof(mine));
    words[8] = bytes;
    CK(cudaMemcpyAsync(c->d_send, words, sizeof(words), cudaMemcpyHostToDevice, c->s_ctl));
    NK(g_nccl.AllGather(c->d_send, c->d_recv, 16, ncclUint64, c->ctl, c->s_ctl));
    CK(cudaMemcpyAsync(c->h_recv, c->d_recv, 8 * 16 * (size_t)c->nranks, cudaMemcpyDeviceToHost, c->s_ctl));
    CK(cudaStreamSynchronize(c->s_ctl));
    for (int i = 0; i < c->nranks; i++) {
        c->sh_bytes[i] = c->h_recv[16 * i + 8];
        if (i == c->rank) { c->sh_peer[i] = c->sh_local; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, c->h_recv + 16 * i, sizeof(h));
        void* ptr = nullptr;
        CK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        c->sh_peer[i] = (uint8_t*)ptr;
    }
    if (d_gathered) *d_gathered = gathered_bytes ? (void*)(c->sh_local + kSharedCtl) : nullptr;
    return SPZ_OK;
}

extern "C" int spz_frame_encode_gather_p2p_dev(spz_comm_t comm, int dst, int predictor, int elem_sz, const void* d_src,
                                               uint64_t total_len, uint16_t ndims, uint32_t chunk_rows, uint32_t stripe_chunks,
                                               void* d_frame, uint64_t frame_cap, void* d_work, uint64_t work_bytes,
                                               spz_stream_t stream_)
{
    Comm* c = (Comm*)comm;
    if (!c || dst < 0 || dst >= c->nranks || stripe_chunks == 0 || ndims == 0 || chunk_rows == 0) return SPZ_E_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    if (!c->sh_local || !c->sh_peer[dst]) return fail_msg(SPZ_E_ARG, "encode_gather_p2p: call spz_comm_shared_alloc first");
    cudaStream_t st = (cudaStream_t)stream_;
    Plan pl;
    if (!make_plan(pl, total_len, ndims, chunk_rows, stripe_chunks, c->nranks)) return SPZ_E_UNSUPPORTED;
    const FrameLayout& gf = pl.f;
    if (c->sh_bytes[dst] < kSharedCtl + spz_frame_max_bytes(elem_sz, total_len, ndims, chunk_rows)) return SPZ_E_CAPACITY;
    const uint64_t len = plan_local_len(pl, c->rank, total_len);
    if (len == 0) return fail_msg(SPZ_E_ARG, "encode_gather_p2p: every rank must own at least one stripe");
    FrameJob j;
    int rc = frame_job_init(j, predictor, elem_sz, d_src, len, ndims, chunk_rows, d_frame, frame_cap, d_work, work_bytes,
                            nullptr, stripe_chunks);
    if (rc != SPZ_OK) return rc;
    if (j.slab_chunks != stripe_chunks || j.f.nchunks != pl.local_chunks[c->rank]) return SPZ_E_ARG;
    const uint64_t nrounds = (pl.segs.size() + c->nranks - 1) / c->nranks;     // I own a stripe in round k iff k < j.nslabs
    if (nrounds > (uint64_t)kMaxLocalStripes) return SPZ_E_UNSUPPORTED;
    if ((rc = frame_job_reset(j, st)) != SPZ_OK) return rc;
    SidePipe& sp = side_pipe();
    CK(sp.init((int)j.nslabs));
    c->epoch++;
    const unsigned long long tag = (unsigned long long)(c->epoch % 0xfffeu) + 1ull;        // 1 .. 0xfffe, never 0
    PeerTab pt;
    for (int i = 0; i < kMaxRanks; i++) pt.ctl[i] = i < c->nranks ? (SharedCtl*)c->sh_peer[i] : nullptr;
    SharedCtl* me = (SharedCtl*)c->sh_local;
    uint8_t* gframe = c->sh_peer[dst] + kSharedCtl;                     // the gathered frame, in the destination's memory
    const uint64_t pay_cap = c->sh_bytes[dst] - kSharedCtl - gf.payload_off;

    CK(cudaEventRecord(c->ev_main, st));
    CK(cudaStreamWaitEvent(sp.side, c->ev_main, 0));
    CK(cudaStreamWaitEvent(c->s_p2p, c->ev_main, 0));
    CK(cudaStreamWaitEvent(c->s_enc[0], c->ev_main, 0));
    CK(cudaStreamWaitEvent(c->s_enc[1], c->ev_main, 0));
    // A stripe is usually less than one machine round of the codec kernel, and a kernel lasts at least one CTA lifetime:
    // back to back on ONE stream eight stripes of 0.43 rounds take eight CTA lifetimes.  On three streams in turn the next
    // stripe's CTAs fill the machine as the previous one's retire.
    cudaStream_t es[3] = {st, c->s_enc[0], c->s_enc[1]};
    CK(cudaMemsetAsync(&me->err, 0, sizeof(unsigned int), c->s_p2p));
    const uint64_t* l_off = (const uint64_t*)((uint8_t*)d_frame + j.f.off_dir);
    const uint32_t* l_nb = (const uint32_t*)((uint8_t*)d_frame + j.f.off_sizes);
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    for (uint64_t k = 0; k < nrounds; k++) {
        const bool has = k < j.nslabs;
        uint64_t c0 = 0, nc = 0, e0 = 0, ne = 0;
        if (has) {
            j.slab(k, c0, nc, e0, ne);
            if ((rc = frame_job_encode_slab(j, k, es[k % 3])) != SPZ_OK) return rc;
            CK(cudaEventRecord(sp.ev_enc[k], es[k % 3]));
            CK(cudaStreamWaitEvent(sp.side, sp.ev_enc[k], 0));
            if ((rc = frame_job_scan_slab(j, k, sp.side)) != SPZ_OK) return rc;   // sizes -> local offsets, running total
        }
        count_launch();
        spz_p2p_publish_kernel<<<1, 64, 0, sp.side>>>(pt, c->nranks, c->rank, (int)k, j.d_tot(), has ? 1 : 0, tag);
        CK(cudaGetLastError());
        CK(cudaEventRecord(c->ev_c[k], sp.side));
        CK(cudaStreamWaitEvent(c->s_p2p, c->ev_c[k], 0));
        count_launch();
        spz_p2p_place_kernel<<<1, 64, 0, c->s_p2p>>>(me, c->nranks, c->rank, (int)k, tag);
        CK(cudaGetLastError());
        if (has) {
            const uint64_t g_lo = (k * (uint64_t)c->nranks + (uint64_t)c->rank) * stripe_chunks;      // first global chunk of the stripe
            uint64_t grid = (nc + 7) / 8;
            if (grid > (uint64_t)sms) grid = (uint64_t)sms;             // one CTA per SM at most: the encoder keeps the rest
            count_launch();
            spz_p2p_gather_kernel<<<(unsigned)grid, 256, 0, c->s_p2p>>>(
                j.work + j.w.off_slots + c0 * j.w.slot_stride, j.w.slot_stride, nc, l_off + c0, l_nb + c0, j.d_tot() + k, me, (int)k,
                gframe + gf.payload_off, pay_cap, (uint64_t*)(gframe + gf.off_dir) + g_lo, (uint32_t*)(gframe + gf.off_sizes) + g_lo);
            CK(cudaGetLastError());
        }
    }
    count_launch();
    spz_p2p_signal_kernel<<<1, 1, 0, c->s_p2p>>>(me, j.d_err(), (SharedCtl*)c->sh_peer[dst], c->rank, tag);
    CK(cudaGetLastError());
    if (c->rank == dst) {
        spz_frame_header h = j.h;
        h.len = total_len; h.nchunks = gf.nchunks; h.payload_off = gf.payload_off;
        count_launch();
        spz_p2p_finish_kernel<<<1, 64, 0, c->s_p2p>>>(me, c->nranks, (int)nrounds, tag, (uint64_t*)(gframe + gf.off_dir), gf.nchunks,
                                                     (uint32_t*)(gframe + gf.off_sizes), (spz_frame_header*)gframe, h);
        CK(cudaGetLastError());
    }
    CK(cudaEventRecord(c->ev_data, c->s_p2p));
    CK(cudaStreamWaitEvent(st, c->ev_data, 0));
    return SPZ_OK;
}
