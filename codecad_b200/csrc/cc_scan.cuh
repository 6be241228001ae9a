// Decoupled look-back across CTAs (the device-wide exclusive prefix used by every ordered
// compaction: classification lists, mass-property lists, marching-cubes triangles).
#ifndef CC_SCAN_CUH
#define CC_SCAN_CUH

#include "cc_device_types.h"

#ifndef CC_DEV
#define CC_DEV __device__ __forceinline__
#endif

// ---- ordered compaction: warp-ballot scan inside the CTA + decoupled look-back across CTAs ----
// tile_status word = (state << 62) | value, state 1 = tile aggregate, 2 = inclusive prefix.
#define CC_ST_AGG 1ull
#define CC_ST_INC 2ull

CC_DEV unsigned long long cc_ld_status(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
CC_DEV void cc_st_status(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Returns the exclusive prefix (number of hits in all earlier tiles, plus the list's initial
// length) for this tile.  Called by warp 0 only; `tile` is the ticket-ordered tile id.
CC_DEV uint32_t cc_lookback(unsigned long long *status, uint32_t tile, uint32_t aggregate,
                            const uint32_t *counter)
{
    const uint32_t lane = threadIdx.x & 31;
    if (tile == 0) {
        uint32_t base = *counter;  // list[atomic_inc(counter)]: continue after existing entries
        if (lane == 0) cc_st_status(status, (CC_ST_INC << 62) | (unsigned long long)(base + aggregate));
        return base;
    }
    if (lane == 0) cc_st_status(status + tile, (CC_ST_AGG << 62) | (unsigned long long)aggregate);
    uint32_t exclusive = 0;
    int look = (int)tile - 1;  // lanes inspect tiles look - lane
    for (;;) {
        int t = look - (int)lane;
        unsigned long long s = (t >= 0) ? cc_ld_status(status + t) : ((CC_ST_INC << 62));
        uint32_t state = (uint32_t)(s >> 62);
        // all inspected predecessors must be published before we can use the window
        if (__any_sync(0xffffffffu, state == 0)) continue;
        uint32_t inc_mask = __ballot_sync(0xffffffffu, state == CC_ST_INC);
        uint32_t val = (uint32_t)s;
        if (inc_mask) {
            int first = __ffs(inc_mask) - 1;  // nearest tile with an inclusive prefix
            uint32_t contrib = (lane <= (uint32_t)first) ? val : 0u;
            exclusive += __reduce_add_sync(0xffffffffu, contrib);
            break;
        }
        exclusive += __reduce_add_sync(0xffffffffu, val);
        look -= 32;
    }
    if (lane == 0)
        cc_st_status(status + tile, (CC_ST_INC << 62) | (unsigned long long)(exclusive + aggregate));
    return exclusive;
}

#endif
