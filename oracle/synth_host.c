/* oracle/synth_host.c — multi-threaded host-side generator for the synthetic assets of include/lt_synth.h.
 * Bench/test infrastructure: fills the host buffers the CPU baseline (the unmodified reference) reads. */
#include "../include/lt_synth.h"

#include <pthread.h>
#include <stdlib.h>

struct job
{
    struct lt_synth_spec spec;
    uint64_t asset_id, offset, len;
    uint8_t* dst;
};

static void* run(void* p)
{
    struct job* j = (struct job*)p;
    lt_synth_fill(&j->spec, j->asset_id, j->offset, j->dst, j->len);
    return 0;
}

__attribute__((visibility("default"))) int synth_fill_mt(const struct lt_synth_spec* spec, uint64_t asset_id, uint64_t offset,
                                                          uint8_t* dst, uint64_t len, uint32_t threads)
{
    if (threads == 0) threads = 1;
    if (threads > 256) threads = 256;
    struct job jobs[256];
    pthread_t tids[256];
    uint64_t per = ((len / threads) + (LT_SYNTH_SEGMENT_BYTES - 1)) & ~(uint64_t)(LT_SYNTH_SEGMENT_BYTES - 1);
    if (per == 0) per = LT_SYNTH_SEGMENT_BYTES;
    uint32_t n = 0;
    for (uint64_t o = 0; o < len && n < threads; o += per, ++n)
    {
        jobs[n].spec = *spec;
        jobs[n].asset_id = asset_id;
        jobs[n].offset = offset + o;
        jobs[n].dst = dst + o;
        jobs[n].len = (n == threads - 1 || o + per > len) ? len - o : per;
        pthread_create(&tids[n], 0, run, &jobs[n]);
        if (n == threads - 1) { ++n; break; }
    }
    for (uint32_t i = 0; i < n; ++i) pthread_join(tids[i], 0);
    return 0;
}
