/* lt_lz4_segments.c — TEST INFRASTRUCTURE: a CPU model of the speculative-segment LZ4 encoder planned for k_lz4_blocks (DESIGN.md section 8,
 * item 1), written to check its one non-obvious claim before any kernel is built on it:
 *
 *   the future of LZ4_compress_generic (lib/lz4/ext/lz4.c:930-1338, byU32 mode) from a point right after a match was emitted depends only on
 *   (position, hash-table entries that are at most 65 535 bytes old) — entries further back are rejected by the distance check (:1089, :1262)
 *   and the search state (step, searchMatchNb) is reset by every match (:1033-1036).
 *
 * Model: worker 0 parses from the block start.  Worker k > 0 starts `warm` bytes before its segment start H_k with an EMPTY table, parses
 * without emitting until its first sequence boundary >= H_k and keeps (boundary, table) as its speculative state S_k.  When the running parse
 * reaches ITS first boundary >= H_k it compares: same boundary, and every slot equal or both entries older than 65 535 bytes.  If so, parsing
 * continues from S_k (on the GPU: worker k already did that in parallel), else it just keeps going.  The output must equal
 * lto_lz4_compress's for every input, whatever the handover decisions were; tests/test_oracle_lz4_segments.py checks that and reports how
 * often handovers are accepted on the benchmark's data classes.
 *
 * The parser below is the one of lt_oracle.c (same citations), made resumable at sequence boundaries. */
#include <errno.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define TABLE_ENTRIES 4096u

static inline uint32_t load32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t load64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline uint32_t hash5(const uint8_t* p) { return (uint32_t)(((load64(p) << 24) * 889523592379ull) >> 52); } /* lz4.c:785-795 */

static uint8_t* put_length(uint8_t* op, uint64_t len)
{
    for (; len >= 255; len -= 255) *op++ = 255;
    *op++ = (uint8_t)len;
    return op;
}

typedef struct
{
    uint64_t ip;     /* == anchor: a match just ended here (or: the block start for worker 0, see `fresh`) */
    uint32_t table[TABLE_ENTRIES];
} seg_state;

/* Runs the parser from `st`.  fresh: the block start (lz4.c:1004-1010) instead of a sequence boundary.
 * Stops at the first sequence boundary >= stop_at (returns 1, st updated, nothing of the next sequence emitted) or at the end of the
 * block (returns 0 after the last literals).  op may be NULL (speculative warm-up: nothing is emitted). */
static int seg_run(const uint8_t* src, uint64_t n, seg_state* st, int fresh, uint64_t stop_at, uint8_t** op_io)
{
    uint32_t* table = st->table;
    uint8_t* op = op_io ? *op_io : 0;
    const uint64_t mflimit_plus_one = n - 11, matchlimit = n - 5;
    uint64_t ip = st->ip, anchor = st->ip;
    uint32_t forward_h = 0;
    uint64_t match = 0;
    uint8_t* token = 0;
    uint8_t scratch_token = 0;
    int resumed = !fresh;
    if (fresh)
    {
        table[hash5(src)] = 0;
        ++ip;
        forward_h = hash5(src + ip);
    }
    for (;;)
    {
        if (resumed)
        {
            resumed = 0;
            goto boundary; /* a match just ended at ip == anchor */
        }
        { /* lz4.c:1043-1100 */
            uint64_t forward_ip = ip;
            uint32_t step = 1, search_nb = 64;
            for (;;)
            {
                uint32_t h = forward_h;
                uint64_t current = forward_ip;
                uint64_t match_index = table[h];
                ip = forward_ip;
                forward_ip += step;
                step = search_nb++ >> 6;
                if (forward_ip > mflimit_plus_one) goto last_literals;
                forward_h = hash5(src + forward_ip);
                table[h] = (uint32_t)current;
                if (match_index + 65535 < current) continue;
                if (load32(src + match_index) == load32(src + ip)) { match = match_index; break; }
            }
        }
        while (ip > anchor && match > 0 && src[ip - 1] == src[match - 1]) { --ip; --match; } /* lz4.c:1104-1109 */
        { /* lz4.c:1112-1137 */
            uint64_t lit = ip - anchor;
            if (op)
            {
                token = op++;
                if (lit >= 15) { *token = 0xF0; op = put_length(op, lit - 15); }
                else *token = (uint8_t)(lit << 4);
                memcpy(op, src + anchor, lit);
                op += lit;
            }
            else token = &scratch_token;
        }
    next_match:
        { /* lz4.c:1157-1226 */
            uint64_t off = ip - match;
            uint64_t a = ip + 4, b = match + 4;
            while (a < matchlimit && src[a] == src[b]) { ++a; ++b; }
            uint64_t code = a - (ip + 4);
            ip = a;
            if (op)
            {
                *op++ = (uint8_t)off;
                *op++ = (uint8_t)(off >> 8);
                if (code >= 15) { *token += 15; op = put_length(op, code - 15); }
                else *token += (uint8_t)code;
            }
        }
        anchor = ip;
        if (ip >= mflimit_plus_one) break; /* lz4.c:1233 */
        if (ip >= stop_at)
        {
            st->ip = ip; /* the hand-over point: nothing of what follows has touched the table yet */
            if (op_io) *op_io = op;
            return 1;
        }
    boundary:
        table[hash5(src + ip - 2)] = (uint32_t)(ip - 2); /* lz4.c:1236-1243 */
        { /* lz4.c:1256-1294 */
            uint32_t h = hash5(src + ip);
            uint64_t match_index = table[h];
            table[h] = (uint32_t)ip;
            if (match_index + 65535 >= ip && load32(src + match_index) == load32(src + ip))
            {
                if (op) { token = op++; *token = 0; }
                else token = &scratch_token;
                match = match_index;
                goto next_match;
            }
        }
        forward_h = hash5(src + ++ip); /* lz4.c:1298 */
    }
last_literals:
    if (op)
    { /* lz4.c:1302-1329 */
        uint64_t last = n - anchor;
        if (last >= 15) { *op++ = 0xF0; op = put_length(op, last - 15); }
        else *op++ = (uint8_t)(last << 4);
        memcpy(op, src + anchor, last);
        op += last;
    }
    if (op_io) *op_io = op;
    return 0;
}

/* every slot equal, or both entries more than 65 535 bytes behind `at` (unobservable from there on) */
static int tables_equivalent(const uint32_t* a, const uint32_t* b, uint64_t at)
{
    for (uint32_t i = 0; i < TABLE_ENTRIES; ++i)
        if (a[i] != b[i] && !((uint64_t)a[i] + 65535 < at && (uint64_t)b[i] + 65535 < at)) return 0;
    return 1;
}

/* segments = K equal slices of the block; out_stats[0] = handovers attempted, [1] = accepted, [2] = rejected because the boundaries differ,
 * [3] = rejected because the tables differ, [4] = bytes parsed speculatively (warm-ups) */
int lto_lz4_compress_segments(const uint8_t* src, uint64_t size, uint32_t segments, uint64_t warm, uint8_t* dst, uint64_t cap, uint64_t* out_size,
                              uint64_t out_stats[5])
{
    memset(out_stats, 0, sizeof(uint64_t) * 5);
    if (size < 65547 || size > 0x7E000000u || cap < size + size / 255 + 16 || segments == 0) return EINVAL; /* byU32 blocks only (lz4.c:710) */
    seg_state* cur = (seg_state*)calloc(1, sizeof(seg_state));
    seg_state* spec = (seg_state*)calloc(1, sizeof(seg_state));
    if (!cur || !spec) { free(cur); free(spec); return ENOMEM; }
    uint8_t* op = dst;
    int fresh = 1, more = 1;
    for (uint32_t k = 1; more; ++k)
    {
        const uint64_t h_k = k < segments ? size / segments * k : (uint64_t)-1; /* segment start; the last worker runs to the end */
        more = seg_run(src, size, cur, fresh, h_k, &op);
        fresh = 0;
        if (!more) break;
        /* worker k's speculation: cold table, `warm` bytes before its segment (it needs 65 536 bytes of history for an empty table to be a
         * legal state: every entry of a zeroed table is out of reach beyond that) */
        if (h_k < warm + 65536) continue;
        memset(spec, 0, sizeof(*spec));
        spec->ip = h_k - warm;
        out_stats[0]++;
        out_stats[4] += warm;
        if (!seg_run(src, size, spec, 0, h_k, 0)) continue; /* ran into the end of the block: nothing to hand over */
        if (spec->ip != cur->ip) { out_stats[2]++; continue; }
        if (!tables_equivalent(spec->table, cur->table, cur->ip)) { out_stats[3]++; continue; }
        out_stats[1]++;
        memcpy(cur, spec, sizeof(*cur)); /* the parse continues from the speculative worker's state */
    }
    *out_size = (uint64_t)(op - dst);
    free(cur);
    free(spec);
    return 0;
}
