// afan_p2p.cuh -- the peer-memory statistics exchange shared by the BatchNorm kernels (afan_bn.cu) and the folded
// BatchNorm of the tcgen05 convolution (afan_conv_umma.cu): mailbox layout, LL-style self-validating words, parameters.
#pragma once
#include <cstdlib>

#include "afan_common.cuh"

namespace afan {

constexpr int kP2PMaxWorld = 8;
constexpr int kP2PRing = 4;
// A lost peer must neither hang the GPU nor go unnoticed: after `timeout_cycles` (AFAN_P2P_TIMEOUT_S, default 60 s --
// the bound on tolerated inter-rank skew, e.g. rank 0 writing a checkpoint) the waiting rank sets state[2] AND poisons
// the folded statistics with NaN, so every later loss on that rank is NaN instead of silently wrong.
constexpr double kP2PDefaultTimeoutS = 60.0;
constexpr double kP2PCyclesPerSecond = 1.9e9;

struct P2PParams {
    void* peers[kP2PMaxWorld];        // peer-mapped mailbox base of every rank (peers[rank] = own mailbox)
    unsigned long long* state;        // local: {seq, ticket, error}
    int world, rank;
    unsigned int cmax;
    long long timeout_cycles;
};

// LL-style in-band flags (the idea of NCCL's low-latency protocol): every double travels as one 16-byte word
// {lo32, tag, hi32, tag}; each 8-byte half carries its own tag, so the word is self-validating however the
// fabric splits the store -- no fence, no separate flag, ONE one-way NVLink latency per exchange.
// word index: ((((slot * world + src) * cmax + ch) * 2 + g) * 2 + k),  k = 0: first sum, 1: second sum.
__host__ __device__ inline size_t p2p_word_off(unsigned int slot, unsigned int src, unsigned int ch, unsigned int g, unsigned int k,
                                               int world, unsigned int cmax) {
    return ((((static_cast<size_t>(slot) * world + src) * cmax + ch) * 2 + g) * 2 + k) * sizeof(uint4);
}
__host__ inline int64_t p2p_mailbox_bytes(int world, int64_t cmax) {
    return static_cast<int64_t>(kP2PRing) * world * cmax * 4 * sizeof(uint4);
}
__device__ __forceinline__ void st_sys_u32x4(uint4* p, uint4 v) {
    asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_sys_u32x4(const uint4* p) {
    uint4 v;
    asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

// the sum travelling as one LL word / the bounded spin that collects it (timeout: error word + NaN, never a hang)
__device__ __forceinline__ void p2p_publish(const P2PParams& q, int peer, unsigned int slot, unsigned int ch, unsigned int g, unsigned int k,
                                            unsigned int tag, double v) {
    const unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(v));
    uint4 w;
    w.x = static_cast<unsigned int>(bits); w.y = tag; w.z = static_cast<unsigned int>(bits >> 32); w.w = tag;
    st_sys_u32x4(reinterpret_cast<uint4*>(static_cast<char*>(q.peers[peer]) + p2p_word_off(slot, q.rank, ch, g, k, q.world, q.cmax)), w);
}
__device__ __forceinline__ double p2p_collect(const P2PParams& q, int src, unsigned int slot, unsigned int ch, unsigned int g, unsigned int k,
                                              unsigned int tag) {
    const uint4* p = reinterpret_cast<const uint4*>(static_cast<const char*>(q.peers[q.rank]) + p2p_word_off(slot, src, ch, g, k, q.world, q.cmax));
    const long long t0 = clock64();
    uint4 r = ld_sys_u32x4(p);
    while (r.y != tag || r.w != tag) {
        if (clock64() - t0 > q.timeout_cycles) {                                  // peer lost: never hang, never pass silently
            q.state[2] = 1ULL;
            r.x = 0u; r.z = 0x7ff80000u;                                          // quiet NaN poisons the statistics
            break;
        }
        r = ld_sys_u32x4(p);
    }
    return __longlong_as_double(static_cast<long long>((static_cast<unsigned long long>(r.z) << 32) | r.x));
}
__device__ __forceinline__ unsigned int p2p_tag(unsigned long long seq) { return static_cast<unsigned int>(seq % 0xfffffffeULL) + 1u; }   // never 0

static int fill_p2p(P2PParams& q, int world, int rank, void* const* peer_mailboxes, int64_t cmax, void* state, int64_t c) {
    if (world < 2 || world > kP2PMaxWorld || rank < 0 || rank >= world || cmax < c) return AFAN_ERR_UNSUPPORTED;
    if (!peer_mailboxes || !state) return AFAN_ERR_NULL;
    for (int i = 0; i < world; ++i) {
        if (!peer_mailboxes[i]) return AFAN_ERR_NULL;
        q.peers[i] = peer_mailboxes[i];
    }
    q.state = static_cast<unsigned long long*>(state);
    q.world = world; q.rank = rank; q.cmax = static_cast<unsigned int>(cmax);
    static const long long timeout = [] {
        const char* e = std::getenv("AFAN_P2P_TIMEOUT_S");
        double sec = e ? std::atof(e) : kP2PDefaultTimeoutS;
        if (!(sec > 0.0)) sec = kP2PDefaultTimeoutS;
        return static_cast<long long>(sec * kP2PCyclesPerSecond);
    }();
    q.timeout_cycles = timeout;
    return AFAN_OK;
}

}  // namespace afan
