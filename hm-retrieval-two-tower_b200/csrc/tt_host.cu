// tt_host.cu -- host-side native helpers on either side of the GPU path (SURVEY.md 8f "next" rows 2 and 3).
// No device code here; it lives in libtt.so so that the Python layer binds ONE library.
//
//   * vocabulary map: StringLookup(num_oov_indices=1, vocabulary=v) (reference input_layer.py:33-36, vocab order from
//     features.py:119-127): string -> row id (0 = OOV, v[i] -> i + 1) as an open-addressing hash table with batched,
//     multi-threaded lookups -- the per-example Python dict lookup was the first bottleneck once the kernels were fast;
//   * TFRecord framing (reference tfrecord_writer.py:112-126 writes them with tf.io.TFRecordWriter, tfrecord_dataset.py:
//     86-98 reads them with tf.data.TFRecordDataset): each record is {u64 length, u32 masked CRC32C(length), bytes,
//     u32 masked CRC32C(bytes)}; CRC32C (Castagnoli) with the SSE4.2 instruction when the host has it;
//   * tf.train.Example batch parser: Example{1: Features{1: map<string, Feature>}}, Feature{1: BytesList | 2: FloatList |
//     3: Int64List} with one value per feature (tfrecord_dataset.py:33-36: FixedLenFeature([1], dtype)).
#include <stdint.h>
#include <string.h>

#include <string>
#include <thread>
#include <vector>

#include "tt_common.cuh"

#if defined(__x86_64__)
#include <cpuid.h>
#include <nmmintrin.h>
#endif

namespace tt {
namespace host {

// ---- CRC32C -----------------------------------------------------------------------------------------------------
static uint32_t g_crc_table[8][256];
static bool g_crc_ready = false;
static bool g_crc_hw = false;

static void crc_init() {
    if (g_crc_ready) return;
    for (uint32_t i = 0; i < 256; ++i) {
        uint32_t c = i;
        for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : (c >> 1);
        g_crc_table[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
        for (int t = 1; t < 8; ++t) g_crc_table[t][i] = (g_crc_table[t - 1][i] >> 8) ^ g_crc_table[0][g_crc_table[t - 1][i] & 0xFF];
#if defined(__x86_64__)
    unsigned a, b, c, d;
    if (__get_cpuid(1, &a, &b, &c, &d)) g_crc_hw = (c & (1u << 20)) != 0;   // SSE4.2
#endif
    g_crc_ready = true;
}

#if defined(__x86_64__)
__attribute__((target("sse4.2"))) static uint32_t crc_hw(uint32_t crc, const unsigned char* p, size_t n) {
    uint64_t c = crc;
    while (n >= 8) {
        uint64_t v;
        memcpy(&v, p, 8);
        c = _mm_crc32_u64(c, v);
        p += 8; n -= 8;
    }
    uint32_t c32 = (uint32_t)c;
    while (n--) c32 = _mm_crc32_u8(c32, *p++);
    return c32;
}
#endif

static uint32_t crc_sw(uint32_t crc, const unsigned char* p, size_t n) {
    while (n >= 8) {
        uint32_t lo, hi;
        memcpy(&lo, p, 4);
        memcpy(&hi, p + 4, 4);
        lo ^= crc;
        crc = g_crc_table[7][lo & 0xFF] ^ g_crc_table[6][(lo >> 8) & 0xFF] ^ g_crc_table[5][(lo >> 16) & 0xFF] ^ g_crc_table[4][lo >> 24] ^
              g_crc_table[3][hi & 0xFF] ^ g_crc_table[2][(hi >> 8) & 0xFF] ^ g_crc_table[1][(hi >> 16) & 0xFF] ^ g_crc_table[0][hi >> 24];
        p += 8; n -= 8;
    }
    while (n--) crc = (crc >> 8) ^ g_crc_table[0][(crc ^ *p++) & 0xFF];
    return crc;
}

static uint32_t crc32c(const void* data, size_t n, bool force_sw = false) {
    crc_init();
    const unsigned char* p = reinterpret_cast<const unsigned char*>(data);
    uint32_t crc = 0xFFFFFFFFu;
#if defined(__x86_64__)
    if (g_crc_hw && !force_sw) return crc_hw(crc, p, n) ^ 0xFFFFFFFFu;
#endif
    return crc_sw(crc, p, n) ^ 0xFFFFFFFFu;
}
static inline uint32_t mask_crc(uint32_t crc) { return ((crc >> 15) | (crc << 17)) + 0xa282ead8u; }   // TFRecord's masking

// ---- vocabulary map -----------------------------------------------------------------------------------------------
// Word-at-a-time multiply / xor-shift hash.  Only speed depends on it: every hit is confirmed by comparing the bytes.
static inline uint64_t hash_bytes(const char* p, size_t n) {
    uint64_t h = 0x9e3779b97f4a7c15ull ^ (n * 0xd6e8feb86659fd93ull);
    while (n >= 8) {
        uint64_t w;
        memcpy(&w, p, 8);
        h = (h ^ w) * 0xff51afd7ed558ccdull;
        h ^= h >> 29;
        p += 8; n -= 8;
    }
    if (n) {
        uint64_t w = 0;
        memcpy(&w, p, n);
        h = (h ^ w) * 0xc4ceb9fe1a85ec53ull;
    }
    h ^= h >> 32; h *= 0xd6e8feb86659fd93ull; h ^= h >> 32;
    return h;
}

// Open addressing, load factor <= 1/2.  An entry carries everything a probe needs before it touches the string bytes -- the high
// hash bits, the vocabulary index and where the string lies -- so a lookup costs two dependent cache misses (entry, bytes), and
// lookup_block overlaps those misses across 16 keys with software prefetches (a 1.37 M-entry table does not fit any cache).
struct Vocab {
    struct Ent {
        uint32_t tag;       // high hash bits | 1; 0 = empty slot
        int32_t idx1;       // vocabulary index + 1
        uint64_t where;     // (offset into blob << 24) | length
    };
    static constexpr uint64_t kMaxLen = (1ull << 24) - 1;
    std::string blob;       // all vocabulary strings back to back
    std::vector<Ent> ent;
    uint64_t mask = 0;
    int64_t n = 0;

    static inline uint32_t tag_of(uint64_t h) { return (uint32_t)(h >> 32) | 1u; }

    // first slot at or after `i` that is empty or carries the tag
    inline uint64_t seek(uint64_t i, uint32_t tg) const {
        while (ent[i].tag != 0 && ent[i].tag != tg) i = (i + 1) & mask;
        return i;
    }
    // resolve from a slot returned by seek
    inline int32_t finish(uint64_t i, uint32_t tg, const char* s, size_t len) const {
        for (;; i = (i + 1) & mask) {
            const Ent& e = ent[i];
            if (e.tag == 0) return 0;   // OOV
            if (e.tag == tg && (e.where & kMaxLen) == len && memcmp(blob.data() + (e.where >> 24), s, len) == 0) return e.idx1;
        }
    }
    int32_t find(const char* s, size_t len) const {
        const uint64_t h = hash_bytes(s, len);
        const uint32_t tg = tag_of(h);
        return finish(seek(h & mask, tg), tg, s, len);
    }
    // keys lo..hi-1 through key_at(i, &ptr, &len), results to out[i]
    template <typename KeyAt>
    void lookup_block(int64_t lo, int64_t hi, KeyAt key_at, int32_t* out) const {
        constexpr int W = 16;
        const char* s[W];
        size_t len[W];
        uint64_t pos[W];
        uint32_t tg[W];
        for (int64_t i = lo; i < hi; i += W) {
            const int m = (int)(hi - i < W ? hi - i : W);
            for (int k = 0; k < m; ++k) {
                key_at(i + k, &s[k], &len[k]);
                const uint64_t h = hash_bytes(s[k], len[k]);
                tg[k] = tag_of(h);
                pos[k] = h & mask;
                __builtin_prefetch(&ent[pos[k]]);
            }
            for (int k = 0; k < m; ++k) {
                pos[k] = seek(pos[k], tg[k]);
                if (ent[pos[k]].tag != 0) __builtin_prefetch(blob.data() + (ent[pos[k]].where >> 24));
            }
            for (int k = 0; k < m; ++k) out[i + k] = finish(pos[k], tg[k], s[k], len[k]);
        }
    }
};

template <typename F>
static void parallel_for(int64_t n, int nthreads, F fn) {
    if (nthreads <= 1 || n < 4096) { fn(0, n); return; }
    std::vector<std::thread> th;
    const int64_t per = (n + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; ++t) {
        const int64_t lo = t * per, hi = lo + per < n ? lo + per : n;
        if (lo >= hi) break;
        th.emplace_back([=]() { fn(lo, hi); });
    }
    for (auto& x : th) x.join();
}

// ---- protobuf wire format (just enough for tf.train.Example) --------------------------------------------------------
struct Cur {
    const unsigned char* p;
    const unsigned char* end;
    bool ok = true;
    uint64_t varint() {
        uint64_t v = 0;
        for (int s = 0; s < 64 && p < end; s += 7) {
            const unsigned char b = *p++;
            v |= (uint64_t)(b & 0x7F) << s;
            if (!(b & 0x80)) return v;
        }
        ok = false;
        return 0;
    }
    Cur sub() {   // length-delimited field body
        const uint64_t n = varint();
        Cur c{p, p + n};
        if (!ok || n > (uint64_t)(end - p)) { ok = false; c.end = c.p; return c; }
        p += n;
        return c;
    }
    void skip(uint32_t wire) {
        if (wire == 0) varint();
        else if (wire == 1) p += 8;
        else if (wire == 2) sub();
        else if (wire == 5) p += 4;
        else ok = false;
        if (p > end) ok = false;
    }
};

}  // namespace host
}  // namespace tt

using namespace tt;
using namespace tt::host;

extern "C" {

uint32_t tt_crc32c(const void* data, size_t n) { return crc32c(data, n); }
uint32_t tt_crc32c_portable(const void* data, size_t n) { return crc32c(data, n, true); }   // table path (tests pin it against the SSE4.2 one)
uint32_t tt_crc32c_masked(const void* data, size_t n) { return mask_crc(crc32c(data, n)); }

void* tt_vocab_create(const char* blob, const int64_t* offsets, int64_t n) {
    if (n < 0 || (n > 0 && (!blob || !offsets))) { set_error("tt_vocab_create: null pointer"); return nullptr; }
    for (int64_t i = 0; i < n; ++i) {
        if (offsets[i + 1] < offsets[i] || (uint64_t)(offsets[i + 1] - offsets[i]) > Vocab::kMaxLen) {
            set_error("tt_vocab_create: entry %lld has a negative or oversized length", (long long)i);
            return nullptr;
        }
    }
    Vocab* v = new Vocab();
    v->n = n;
    v->blob.assign(blob ? blob : "", n > 0 ? (size_t)offsets[n] : 0);
    uint64_t cap = 16;
    while (cap < (uint64_t)n * 2 + 2) cap <<= 1;
    v->ent.assign(cap, Vocab::Ent{0u, 0, 0ull});
    v->mask = cap - 1;
    for (int64_t i = 0; i < n; ++i) {
        const char* s = v->blob.data() + offsets[i];
        const size_t len = (size_t)(offsets[i + 1] - offsets[i]);
        const uint64_t h = hash_bytes(s, len);
        const uint32_t tg = Vocab::tag_of(h);
        uint64_t j = h & v->mask;
        bool dup = false;
        for (;; j = (j + 1) & v->mask) {
            const Vocab::Ent& e = v->ent[j];
            if (e.tag == 0) break;
            if (e.tag == tg && (e.where & Vocab::kMaxLen) == len && memcmp(v->blob.data() + (e.where >> 24), s, len) == 0) { dup = true; break; }
        }
        if (!dup) v->ent[j] = Vocab::Ent{tg, (int32_t)(i + 1), ((uint64_t)offsets[i] << 24) | (uint64_t)len};   // first occurrence wins
    }
    return v;
}

void tt_vocab_destroy(void* h) { delete reinterpret_cast<Vocab*>(h); }

int64_t tt_vocab_size(void* h) { return h ? reinterpret_cast<Vocab*>(h)->n : -1; }

/* strings = blob[offsets[i] .. offsets[i+1]) */
int tt_vocab_lookup(void* h, const char* blob, const int64_t* offsets, int64_t n, int32_t* out, int nthreads) {
    TT_REQUIRE(h && out && (n == 0 || (blob && offsets)), "tt_vocab_lookup: null pointer");
    const Vocab* v = reinterpret_cast<const Vocab*>(h);
    parallel_for(n, nthreads, [=](int64_t lo, int64_t hi) {
        v->lookup_block(lo, hi, [=](int64_t i, const char** s, size_t* len) { *s = blob + offsets[i]; *len = (size_t)(offsets[i + 1] - offsets[i]); }, out);
    });
    return TT_OK;
}

/* numpy 'S<width>' arrays: n fixed-width cells, NUL padded on the right */
int tt_vocab_lookup_fixed(void* h, const char* data, int64_t n, int width, int32_t* out, int nthreads) {
    TT_REQUIRE(h && out && (n == 0 || data) && width >= 1, "tt_vocab_lookup_fixed: bad argument");
    const Vocab* v = reinterpret_cast<const Vocab*>(h);
    parallel_for(n, nthreads, [=](int64_t lo, int64_t hi) {
        v->lookup_block(lo, hi, [=](int64_t i, const char** s, size_t* len) {
            const char* c = data + i * width;
            size_t l = (size_t)width;
            while (l > 0 && c[l - 1] == '\0') --l;
            *s = c; *len = l;
        }, out);
    });
    return TT_OK;
}

/* Walk the records of one TFRecord file image.  Writes up to `max_records` (offset of the payload, payload length) pairs;
 * returns the number of records found (may exceed max_records: call again with a larger array), or a negative error
 * (TT_ERR_INVALID: truncated file or CRC mismatch; tt_last_error() says where). */
int64_t tt_tfrecord_scan(const void* file, size_t nbytes, int verify_crc, int64_t* rec_offset, int64_t* rec_len, int64_t max_records) {
    const unsigned char* p = reinterpret_cast<const unsigned char*>(file);
    size_t pos = 0;
    int64_t count = 0;
    while (pos < nbytes) {
        if (nbytes - pos < 12) { set_error("tt_tfrecord_scan: truncated header at byte %zu", pos); return TT_ERR_INVALID; }
        uint64_t len;
        uint32_t crc_len;
        memcpy(&len, p + pos, 8);
        memcpy(&crc_len, p + pos + 8, 4);
        if (verify_crc && mask_crc(crc32c(p + pos, 8)) != crc_len) { set_error("tt_tfrecord_scan: length CRC mismatch at byte %zu", pos); return TT_ERR_INVALID; }
        if (len > nbytes - pos - 12 || nbytes - pos - 12 - len < 4) { set_error("tt_tfrecord_scan: truncated record at byte %zu", pos); return TT_ERR_INVALID; }
        uint32_t crc_data;
        memcpy(&crc_data, p + pos + 12 + len, 4);
        if (verify_crc && mask_crc(crc32c(p + pos + 12, (size_t)len)) != crc_data) {
            set_error("tt_tfrecord_scan: data CRC mismatch in record %lld", (long long)count);
            return TT_ERR_INVALID;
        }
        if (count < max_records && rec_offset && rec_len) { rec_offset[count] = (int64_t)(pos + 12); rec_len[count] = (int64_t)len; }
        ++count;
        pos += 12 + (size_t)len + 4;
    }
    return count;
}

/* Frame one payload: out must hold len + 16 bytes. */
int tt_tfrecord_frame(const void* payload, uint64_t len, void* out) {
    TT_REQUIRE(out && (payload || len == 0), "tt_tfrecord_frame: null pointer");
    unsigned char* o = reinterpret_cast<unsigned char*>(out);
    memcpy(o, &len, 8);
    const uint32_t c1 = mask_crc(crc32c(o, 8));
    memcpy(o + 8, &c1, 4);
    if (len) memcpy(o + 12, payload, (size_t)len);
    const uint32_t c2 = mask_crc(crc32c(o + 12, (size_t)len));
    memcpy(o + 12 + len, &c2, 4);
    return TT_OK;
}

/* Parse `nrec` serialized tf.train.Example messages (payload i = file[rec_offset[i] .. +rec_len[i])) for `nfeat` features with
 * ONE value each.  kind[f]: 0 = bytes (string feature), 1 = float.  For a bytes feature f the value of record i is reported as
 * (str_off[f*nrec + i], str_len[f*nrec + i]) into `file`; for a float feature it is written to fvals[f*nrec + i].
 * A record that lacks a requested feature, or holds it with another type or not exactly one value, is an error
 * (tf.io.parse_single_example with FixedLenFeature([1]) raises too). */
int tt_gather_cells(const void* file, size_t nbytes, const int64_t* str_off, const int64_t* str_len, int64_t n, int width, char* out, int nthreads) {
    TT_REQUIRE(out && width >= 1 && n >= 0 && (n == 0 || (file && str_off && str_len)), "tt_gather_cells: bad argument");
    for (int64_t i = 0; i < n; ++i)
        TT_REQUIRE(str_len[i] >= 0 && str_len[i] <= width && str_off[i] >= 0 && (uint64_t)str_off[i] + (uint64_t)str_len[i] <= nbytes,
                   "tt_gather_cells: string %lld lies outside the file image or exceeds the cell width", (long long)i);
    const char* base = reinterpret_cast<const char*>(file);
    parallel_for(n, nthreads, [=](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; ++i) {
            char* cell = out + i * (int64_t)width;
            const size_t len = (size_t)str_len[i];
            memcpy(cell, base + str_off[i], len);
            memset(cell + len, 0, (size_t)width - len);
        }
    });
    return TT_OK;
}

int tt_example_parse(const void* file, const int64_t* rec_offset, const int64_t* rec_len, int64_t nrec, const char* const* names, const int32_t* kind,
                     int nfeat, int64_t* str_off, int64_t* str_len, float* fvals, int nthreads) {
    TT_REQUIRE(file && rec_offset && rec_len && names && kind && nfeat >= 1 && nfeat <= 64, "tt_example_parse: bad argument");
    const unsigned char* base = reinterpret_cast<const unsigned char*>(file);
    std::vector<size_t> name_len(nfeat);
    for (int f = 0; f < nfeat; ++f) name_len[f] = strlen(names[f]);
    std::vector<int64_t> bad(nthreads > 0 ? nthreads : 1, -1);
    int tcount = nthreads > 0 ? nthreads : 1;
    std::vector<std::thread> th;
    const int64_t per = (nrec + tcount - 1) / tcount;
    auto work = [&](int t, int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; ++i) {
            uint64_t seen = 0;
            Cur ex{base + rec_offset[i], base + rec_offset[i] + rec_len[i]};
            while (ex.ok && ex.p < ex.end) {
                const uint64_t key = ex.varint();
                if ((key >> 3) != 1 || (key & 7) != 2) { ex.skip((uint32_t)(key & 7)); continue; }
                Cur feats = ex.sub();   // Features
                while (feats.ok && feats.p < feats.end) {
                    const uint64_t k2 = feats.varint();
                    if ((k2 >> 3) != 1 || (k2 & 7) != 2) { feats.skip((uint32_t)(k2 & 7)); continue; }
                    Cur entry = feats.sub();   // map entry {1: key, 2: Feature}
                    const unsigned char* kname = nullptr;
                    size_t klen = 0;
                    Cur feat{nullptr, nullptr};
                    while (entry.ok && entry.p < entry.end) {
                        const uint64_t k3 = entry.varint();
                        if ((k3 & 7) != 2) { entry.skip((uint32_t)(k3 & 7)); continue; }
                        Cur body = entry.sub();
                        if ((k3 >> 3) == 1) { kname = body.p; klen = (size_t)(body.end - body.p); }
                        else if ((k3 >> 3) == 2) feat = body;
                    }
                    if (!entry.ok || !kname || !feat.p) continue;
                    int f = -1;
                    for (int g = 0; g < nfeat; ++g)
                        if (name_len[g] == klen && memcmp(names[g], kname, klen) == 0) { f = g; break; }
                    if (f < 0) continue;   // a feature nobody asked for
                    // Feature: oneof {1: BytesList{1: repeated bytes}, 2: FloatList{1: packed/repeated float}, 3: Int64List}
                    while (feat.ok && feat.p < feat.end) {
                        const uint64_t k4 = feat.varint();
                        if ((k4 & 7) != 2) { feat.skip((uint32_t)(k4 & 7)); continue; }
                        Cur lst = feat.sub();
                        const int which = (int)(k4 >> 3);
                        if (which == 1 && kind[f] == 0) {
                            int nval = 0;
                            while (lst.ok && lst.p < lst.end) {
                                const uint64_t k5 = lst.varint();
                                if ((k5 >> 3) == 1 && (k5 & 7) == 2) {
                                    Cur val = lst.sub();
                                    str_off[(int64_t)f * nrec + i] = (int64_t)(val.p - base);
                                    str_len[(int64_t)f * nrec + i] = (int64_t)(val.end - val.p);
                                    ++nval;
                                } else lst.skip((uint32_t)(k5 & 7));
                            }
                            if (nval == 1) seen |= 1ull << f;
                        } else if (which == 2 && kind[f] == 1) {
                            int nval = 0;
                            while (lst.ok && lst.p < lst.end) {
                                const uint64_t k5 = lst.varint();
                                if ((k5 >> 3) == 1 && (k5 & 7) == 2) {          // packed floats
                                    Cur pk = lst.sub();
                                    while (pk.p + 4 <= pk.end) { memcpy(&fvals[(int64_t)f * nrec + i], pk.p, 4); pk.p += 4; ++nval; }
                                } else if ((k5 >> 3) == 1 && (k5 & 7) == 5) {   // unpacked float
                                    if (lst.p + 4 <= lst.end) { memcpy(&fvals[(int64_t)f * nrec + i], lst.p, 4); ++nval; }
                                    lst.p += 4;
                                } else lst.skip((uint32_t)(k5 & 7));
                            }
                            if (nval == 1) seen |= 1ull << f;
                        }
                    }
                }
                if (!feats.ok) ex.ok = false;
            }
            const uint64_t want = nfeat == 64 ? ~0ull : ((1ull << nfeat) - 1);
            if (!ex.ok || seen != want) { if (bad[t] < 0) bad[t] = i; }
        }
    };
    if (tcount <= 1 || nrec < 1024) work(0, 0, nrec);
    else {
        for (int t = 0; t < tcount; ++t) {
            const int64_t lo = t * per, hi = lo + per < nrec ? lo + per : nrec;
            if (lo >= hi) break;
            th.emplace_back(work, t, lo, hi);
        }
        for (auto& x : th) x.join();
    }
    for (int t = 0; t < tcount; ++t)
        if (bad[t] >= 0) {
            set_error("tt_example_parse: record %lld is malformed or lacks a requested feature (one value of the declared type each)", (long long)bad[t]);
            return TT_ERR_INVALID;
        }
    return TT_OK;
}

}  // extern "C"
