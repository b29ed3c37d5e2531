/* ORACLE (test infrastructure only -- never linked into the product path).
 *
 * Sequential CPU restatement of MetaFEM's GPU open-addressing hash table:
 *   reference: src/misc/06_GPU_Dict.jl:2-11   (GPU_hash_64_64, Thomas Wang 64-bit)
 *              src/misc/06_GPU_Dict.jl:13     (_DictSize)
 *              src/misc/06_GPU_Dict.jl:45-93  (FEM_Dict_SetID!: grow / re-insert policy)
 *              src/misc/06_GPU_Dict.jl:125-162 (dict_SetID! probe + chain links)
 *              src/misc/06_GPU_Dict.jl:164-188 (dict_GetID chain-only lookup)
 * The reference inserts with racing atomic_cas; this file inserts keys one at a time in
 * array order, which is one legal outcome of the reference (SURVEY.md §8c determinism caveat).
 * All slot IDs are 1-based as in the reference; 0 = empty.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t wang64(uint64_t a) {
    a = ~a + (a << 21);
    a = a ^ (a >> 24);
    a = a + (a << 3) + (a << 8);
    a = a ^ (a >> 14);
    a = a + (a << 2) + (a << 4);
    a = a ^ (a >> 28);
    a = a + (a << 31);
    return a;
}

uint64_t ora_wang64(uint64_t a) { return wang64(a); }

/* update_TruncID, 06_GPU_Dict.jl:124 */
static inline int32_t trunc_id(uint64_t prev, int64_t size) {
    return (int32_t)(((uint32_t)prev) & (uint32_t)(size - 1)) + 1;
}

int64_t ora_dict_size(int64_t x) {
    if ((double)x < 16.0 * 2.0 / 3.0) return 16;
    int64_t need = (int64_t)((3 * x + 1) / 2); /* ceil(1.5 x) for integer x */
    int64_t s = 1;
    while (s < need) s <<= 1;
    return s;
}

/* dict_SetID! executed for keys[0..m) in order. Arrays are length `size`, 1-based IDs stored. */
void ora_dict_set(int64_t size, uint64_t *keys, uint64_t *hashs, int32_t *hinit, int32_t *hprev,
                  int32_t *hnext, int64_t m, const uint64_t *new_keys, int32_t *new_ids) {
    for (int64_t t = 0; t < m; ++t) {
        uint64_t key = new_keys[t];
        uint64_t h = wang64(key);
        int32_t start = trunc_id(h, size);
        int32_t cur = hinit[start - 1] == 0 ? start : hinit[start - 1];
        int32_t last_front = 0;
        for (;;) {
            uint64_t local = keys[cur - 1];
            if (local == 0) {
                keys[cur - 1] = key;
                hashs[cur - 1] = h;
                if (last_front != 0) {
                    hprev[cur - 1] = last_front;
                    hnext[last_front - 1] = cur;
                } else {
                    hinit[start - 1] = cur;
                }
                break;
            } else if (local == key) {
                break;
            } else if (trunc_id(wang64(local), size) == start) {
                if (hnext[cur - 1] == 0) {
                    last_front = cur;
                    cur = trunc_id((uint64_t)cur, size);
                } else {
                    cur = hnext[cur - 1];
                }
            } else {
                cur = trunc_id((uint64_t)cur, size);
            }
        }
        new_ids[t] = cur;
    }
}

/* dict_GetID */
void ora_dict_get(int64_t size, const uint64_t *keys, const uint64_t *hashs, const int32_t *hinit,
                  const int32_t *hnext, int64_t m, const uint64_t *target, int32_t *ids) {
    for (int64_t t = 0; t < m; ++t) {
        uint64_t key = target[t];
        int32_t start = trunc_id(wang64(key), size);
        int32_t cur = hinit[start - 1];
        if (cur == 0) { ids[t] = 0; continue; }
        for (;;) {
            if (keys[cur - 1] == key) { ids[t] = cur; break; }
            else if (trunc_id(hashs[cur - 1], size) != start) { ids[t] = -1; break; }
            else if (hnext[cur - 1] == 0) { ids[t] = 0; break; }
            cur = hnext[cur - 1];
        }
    }
}
