// Host-side ingestion of the worker -> learner wire records.
//
// The reference ships every return as a proto3 `Return` (or a `ReturnArray` of them) over gRPC
// (networking/rpc_misc/proto/client_server_interface.proto:30-47), turns each into an `FDReturn` object
// (learner/fd_return.py:41-56, networking/server.py:151-162) and later walks those objects one by one in the
// learner (learner/finite_differences.py:94-114).  Here the serialized bytes are decoded straight into the
// structure-of-arrays form the device learner uploads (epoch / table index / sign / reward ...), with no
// per-return objects.  Pure host code: no CUDA call, no allocation, the caller owns every array.
//
// Wire format restated from the published proto3 encoding: a message is a sequence of (tag = field << 3 | type)
// varints followed by the value; type 0 = varint, 1 = 8 bytes, 2 = length-delimited, 5 = 4 bytes.  Repeated scalars
// are packed (one length-delimited blob) by every proto3 writer the reference uses; the legal but unusual unpacked
// or split-blob forms make the decoder return DFD_WIRE_UNSUPPORTED so the caller can take its general path.
#include <string.h>

#include "../../include/dfd_b200.h"
#include "common.cuh"

namespace {

struct Cursor {
    const uint8_t* p;
    const uint8_t* end;
};

inline bool read_varint(Cursor& c, uint64_t& v) {
    v = 0;
    for (int shift = 0; shift < 70; shift += 7) {
        if (c.p >= c.end) return false;
        uint8_t b = *c.p++;
        if (shift < 64) v |= (uint64_t)(b & 0x7f) << shift;
        if (!(b & 0x80)) return true;
    }
    return false;
}

inline bool skip_field(Cursor& c, unsigned type) {
    uint64_t n;
    switch (type) {
        case 0: return read_varint(c, n);
        case 1: if (c.end - c.p < 8) return false; c.p += 8; return true;
        case 2: if (!read_varint(c, n) || (uint64_t)(c.end - c.p) < n) return false; c.p += n; return true;
        case 5: if (c.end - c.p < 4) return false; c.p += 4; return true;
        default: return false;   // groups (3, 4) are not proto3
    }
}

inline float load_f32(const uint8_t* p) {
    float f;
    memcpy(&f, p, 4);            // little-endian host (x86-64 / aarch64), as the wire
    return f;
}

// `encoded_noise` of a table source is the decimal index (utils/noise_sources.py:46); '+i' / '-i' is the antithetic
// extension.  Anything else (RNGNoiseSource's "state,inc", an empty key) is left to the caller: idx = -1, sign = 0.
inline void parse_key(const uint8_t* s, int64_t n, int64_t* idx, int8_t* sign) {
    *idx = -1;
    *sign = 0;
    int64_t i = 0;
    int8_t sg = 1;
    if (n > 0 && (s[0] == '+' || s[0] == '-')) { sg = s[0] == '-' ? -1 : 1; i = 1; }
    if (i >= n || n - i > 18) return;
    int64_t v = 0;
    for (; i < n; ++i) {
        if (s[i] < '0' || s[i] > '9') return;
        v = v * 10 + (s[i] - '0');
    }
    *idx = v;
    *sign = sg;
}

// One `Return` message occupying [c.p, c.end).  Returns 0, DFD_WIRE_MALFORMED or DFD_WIRE_UNSUPPORTED.
int decode_one(Cursor c, const uint8_t* base, int64_t j, const dfd_return_soa* o) {
    int64_t epoch = 0, timesteps = 0;
    float reward = 0.f, novelty = 0.f, entropy = 0.f;
    uint8_t is_eval = 0;
    int64_t key_off = 0, states_off = 0, shape_off = 0, stats_off = 0;
    int64_t key_len = 0, states_len = -1, shape_len = -1, stats_len = -1;
    while (c.p < c.end) {
        uint64_t tag, n;
        if (!read_varint(c, tag)) return DFD_WIRE_MALFORMED;
        unsigned field = (unsigned)(tag >> 3), type = (unsigned)(tag & 7);
        if (field >= 8 && field <= 10 && type != 2) return DFD_WIRE_UNSUPPORTED;   // unpacked repeated scalar
        if (type == 2) {
            if (!read_varint(c, n) || (uint64_t)(c.end - c.p) < n) return DFD_WIRE_MALFORMED;
            int64_t off = c.p - base;
            switch (field) {
                case 2: key_off = off; key_len = (int64_t)n; break;
                case 8: if (states_len >= 0) return DFD_WIRE_UNSUPPORTED; states_off = off; states_len = (int64_t)n; break;
                case 9: if (shape_len >= 0) return DFD_WIRE_UNSUPPORTED; shape_off = off; shape_len = (int64_t)n; break;
                case 10: if (stats_len >= 0) return DFD_WIRE_UNSUPPORTED; stats_off = off; stats_len = (int64_t)n; break;
                default: break;
            }
            c.p += n;
        } else if (type == 0 && (field == 1 || field == 6 || field == 7)) {
            if (!read_varint(c, n)) return DFD_WIRE_MALFORMED;
            if (field == 1) epoch = (int64_t)n;
            else if (field == 6) timesteps = (int64_t)(int32_t)(uint32_t)n;
            else is_eval = n != 0;
        } else if (type == 5 && field >= 3 && field <= 5) {
            if (c.end - c.p < 4) return DFD_WIRE_MALFORMED;
            float f = load_f32(c.p);
            c.p += 4;
            if (field == 3) reward = f; else if (field == 4) novelty = f; else entropy = f;
        } else if (!skip_field(c, type)) {
            return DFD_WIRE_MALFORMED;
        }
    }
    if ((states_len > 0 && states_len % 4) || (stats_len > 0 && stats_len % 4)) return DFD_WIRE_MALFORMED;
    o->epoch[j] = epoch;
    parse_key(base + key_off, key_len, &o->idx[j], &o->sign[j]);
    o->reward[j] = (double)reward;     // the wire carries fp32 (proto:33); widened exactly as Python does on read
    o->novelty[j] = novelty;
    o->entropy[j] = entropy;
    o->timesteps[j] = (int32_t)timesteps;
    o->is_eval[j] = is_eval;
    o->key_off[j] = key_off;        o->key_len[j] = (int32_t)key_len;
    o->states_off[j] = states_off;  o->states_len[j] = (int32_t)(states_len < 0 ? 0 : states_len);
    o->shape_off[j] = shape_off;    o->shape_len[j] = (int32_t)(shape_len < 0 ? 0 : shape_len);
    o->stats_off[j] = stats_off;    o->stats_len[j] = (int32_t)(stats_len < 0 ? 0 : stats_len);
    return 0;
}

}  // namespace

extern "C" int64_t dfd_wire_count_returns(const uint8_t* host_buf, size_t len) {
    Cursor c{host_buf, host_buf + len};
    int64_t n_rets = 0;
    while (c.p < c.end) {
        uint64_t tag;
        if (!read_varint(c, tag)) return DFD_WIRE_MALFORMED;
        if ((tag >> 3) == 1 && (tag & 7) == 2) ++n_rets;
        if (!skip_field(c, (unsigned)(tag & 7))) return DFD_WIRE_MALFORMED;
    }
    return n_rets;
}

extern "C" int64_t dfd_wire_decode_returns(const uint8_t* host_buf, size_t len, int is_array, int64_t max_returns,
                                           const dfd_return_soa* out) {
    if (!out || (!host_buf && len)) {
        dfd_set_error("dfd_wire_decode_returns: null argument");
        return DFD_WIRE_MALFORMED;
    }
    if (!is_array) {
        if (max_returns < 1) { dfd_set_error("dfd_wire_decode_returns: output arrays too short"); return DFD_WIRE_MALFORMED; }
        int rc = decode_one(Cursor{host_buf, host_buf + len}, host_buf, 0, out);
        if (rc) { dfd_set_error("dfd_wire_decode_returns: Return message %s", rc == DFD_WIRE_MALFORMED ? "malformed" : "uses an unpacked repeated field"); return rc; }
        return 1;
    }
    Cursor c{host_buf, host_buf + len};
    int64_t j = 0;
    while (c.p < c.end) {
        uint64_t tag, n;
        if (!read_varint(c, tag)) { dfd_set_error("dfd_wire_decode_returns: truncated ReturnArray"); return DFD_WIRE_MALFORMED; }
        if ((tag >> 3) == 1 && (tag & 7) == 2) {
            if (!read_varint(c, n) || (uint64_t)(c.end - c.p) < n) { dfd_set_error("dfd_wire_decode_returns: truncated Return %lld", (long long)j); return DFD_WIRE_MALFORMED; }
            if (j >= max_returns) { dfd_set_error("dfd_wire_decode_returns: more than %lld returns", (long long)max_returns); return DFD_WIRE_MALFORMED; }
            int rc = decode_one(Cursor{c.p, c.p + n}, host_buf, j, out);
            if (rc) { dfd_set_error("dfd_wire_decode_returns: Return %lld %s", (long long)j, rc == DFD_WIRE_MALFORMED ? "malformed" : "uses an unpacked repeated field"); return rc; }
            c.p += n;
            ++j;
        } else if (!skip_field(c, (unsigned)(tag & 7))) {
            dfd_set_error("dfd_wire_decode_returns: malformed ReturnArray");
            return DFD_WIRE_MALFORMED;
        }
    }
    return j;
}
