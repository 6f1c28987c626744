// Host-side streaming decoder: `vg view -j -K` multipath-alignment JSON -> the Reads CSR of include/sfb200.h.
//
// Semantics follow the reference line by line where the result depends on it:
//   grouping and best-score filter   get_all_alignments_one_read   json2csv.py:667-739
//   best routes through the subpath DAG   BFS                      json2csv.py:589-665
//   walk / coverage / error ratios   metaalign2align               json2csv.py:551-587
// The reference re-parses ~30 % of the file in its second pass and runs on one thread.  Here the file is mapped
// and cut into blocks of whole lines; per block
//   A  (parallel over lines)    parse + best routes + walks, into per-thread arenas that are reused (no per-line heap)
//   B  (serial, a few compares per line)   read-group boundaries and the reference's error order
//   C  (parallel over groups)   which lines' alignments a mate keeps (best score, ties), sizes
//   D  (serial prefix sums)     output offsets
//   E  (parallel over groups)   fill the output arrays of the block (allocated once, exact size, not zero-filled)
// A group cut by the end of a block is decoded again at the start of the next one.
#include <algorithm>
#include <chrono>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <zlib.h>
#include <unistd.h>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "sfb200.h"

namespace {

enum Err { E_NONE = 0, E_JSON = 1, E_KEY = 2, E_ZERODIV = 3, E_INDEX = 4, E_IO = 5, E_TYPE = 6 };

struct Fail {
    int kind;
    std::string msg;
};

// ---- a small JSON reader: values are visited in place, only the needed fields are kept -------------
struct Cur {
    const char* p;
    const char* e;
};
inline void ws(Cur& c) {
    while (c.p < c.e && (unsigned char)*c.p <= ' ' && (*c.p == ' ' || *c.p == '\t' || *c.p == '\r' || *c.p == '\n')) ++c.p;
}
[[noreturn]] void bad(const char* what) { throw Fail{E_JSON, std::string("malformed JSON: ") + what}; }

// returns the raw bytes between the quotes (escapes untouched)
inline void str(Cur& c, const char*& b, size_t& n) {
    if (c.p >= c.e || *c.p != '"') bad("expected string");
    b = ++c.p;
    {   // keys and most values are short and hold no escape: scan them in line
        const char* p = b;
        const char* lim = (c.e - b) > 24 ? b + 24 : c.e;
        while (p < lim && *p != '"' && *p != '\\') ++p;
        if (p < lim && *p == '"') { n = (size_t)(p - b); c.p = p + 1; return; }
    }
    for (const char* p = b;;) {
        const char* q = p < c.e ? (const char*)memchr(p, '"', (size_t)(c.e - p)) : nullptr;
        if (!q) bad("unterminated string");
        const char* r = q;                       // an odd run of backslashes in front escapes the quote
        while (r > b && r[-1] == '\\') --r;
        if (((q - r) & 1) == 0) { c.p = q; break; }
        p = q + 1;
    }
    n = (size_t)(c.p - b);
    ++c.p;
}
inline double num(Cur& c) {
    // plain integers (almost every number of a vg alignment) without strtod
    const char* p = c.p;
    bool neg = false;
    if (p < c.e && *p == '-') { neg = true; ++p; }
    if (p < c.e && *p >= '0' && *p <= '9') {
        long long v = 0;
        int digits = 0;
        while (p < c.e && *p >= '0' && *p <= '9' && digits < 18) { v = v * 10 + (*p - '0'); ++p; ++digits; }
        if (p >= c.e || (*p != '.' && *p != 'e' && *p != 'E' && !(*p >= '0' && *p <= '9'))) {
            c.p = p;
            return neg ? -(double)v : (double)v;
        }
    }
    char* end = nullptr;
    const double v = strtod(c.p, &end);
    if (end == c.p) bad("expected number");
    c.p = end;
    return v;
}
void skip(Cur& c);
inline void skip_container(Cur& c, char open, char close) {
    ++c.p;
    ws(c);
    if (c.p < c.e && *c.p == close) { ++c.p; return; }
    for (;;) {
        ws(c);
        if (open == '{') {
            const char* b; size_t n;
            str(c, b, n);
            ws(c);
            if (c.p >= c.e || *c.p != ':') bad("expected ':'");
            ++c.p;
        }
        skip(c);
        ws(c);
        if (c.p >= c.e) bad("unterminated container");
        if (*c.p == ',') { ++c.p; continue; }
        if (*c.p == close) { ++c.p; return; }
        bad("expected ',' or close");
    }
}
void skip(Cur& c) {
    ws(c);
    if (c.p >= c.e) bad("unexpected end");
    switch (*c.p) {
        case '{': skip_container(c, '{', '}'); break;
        case '[': skip_container(c, '[', ']'); break;
        case '"': { const char* b; size_t n; str(c, b, n); break; }
        case 't': c.p += 4; break;
        case 'f': c.p += 5; break;
        case 'n': c.p += 4; break;
        default: num(c);
    }
}
// int(x) of a JSON number or of a string holding one (protobuf JSON writes 64-bit ints as strings)
inline long long int_like(Cur& c) {
    ws(c);
    if (c.p < c.e && *c.p == '"') {
        const char* b; size_t n;
        str(c, b, n);
        if (n >= 1 && n <= 18) {                 // plain digits: the usual case
            long long v = 0;
            size_t i = 0;
            for (; i < n && b[i] >= '0' && b[i] <= '9'; ++i) v = v * 10 + (b[i] - '0');
            if (i == n) return v;
        }
        char* end = nullptr;
        const long long v = strtoll(b, &end, 10);
        if (end != b + n) throw Fail{E_TYPE, "invalid literal for int()"};
        return v;
    }
    return (long long)num(c);
}
template <class F>
inline void object(Cur& c, F f) {
    ws(c);
    if (c.p >= c.e || *c.p != '{') bad("expected object");
    ++c.p;
    ws(c);
    if (c.p < c.e && *c.p == '}') { ++c.p; return; }
    for (;;) {
        ws(c);
        const char* k; size_t n;
        str(c, k, n);
        ws(c);
        if (c.p >= c.e || *c.p != ':') bad("expected ':'");
        ++c.p;
        ws(c);
        if (!f(k, n)) skip(c);
        ws(c);
        if (c.p >= c.e) bad("unterminated object");
        if (*c.p == ',') { ++c.p; continue; }
        if (*c.p == '}') { ++c.p; return; }
        bad("expected ',' or '}'");
    }
}
template <class F>
inline void array(Cur& c, F f) {
    ws(c);
    if (c.p >= c.e || *c.p != '[') bad("expected array");
    ++c.p;
    ws(c);
    if (c.p < c.e && *c.p == ']') { ++c.p; return; }
    for (;;) {
        f();
        ws(c);
        if (c.p >= c.e) bad("unterminated array");
        if (*c.p == ',') { ++c.p; continue; }
        if (*c.p == ']') { ++c.p; return; }
        bad("expected ',' or ']'");
    }
}
inline bool is(const char* k, size_t n, const char* lit) { return n == strlen(lit) && memcmp(k, lit, n) == 0; }

// ---- one alignment line ------------------------------------------------------------------------------
struct Edit {
    int from, to;
    bool has_seq;
};
struct Mapping {
    long long node;
    bool has_edit;
    uint32_t e0, e1;     // range in Parsed::edits
};
struct Sub {
    uint32_t m0, m1;     // range in Parsed::maps
    uint32_t n0, n1;     // range in Parsed::nexts
    double score;
    bool has_next;
};
struct Parsed {          // the fields of one line that the decode needs; per thread, storage reused line after line
    const char* name; size_t name_n;
    const char* pair; size_t pair_n;
    const char* seq; size_t seq_n;
    bool has_seq, has_subpath, has_start;
    std::vector<int> start, nexts;
    std::vector<Sub> subs;
    std::vector<Mapping> maps;
    std::vector<Edit> edits;
    void reset() {
        name = pair = seq = nullptr;
        name_n = pair_n = seq_n = 0;
        has_seq = has_subpath = has_start = false;
        start.clear();
        nexts.clear();
        subs.clear();
        maps.clear();
        edits.clear();
    }
};

void parse_line(const char* b, const char* e, Parsed& ln) {
    Cur c{b, e};
    ln.reset();
    bool has_name = false;
    object(c, [&](const char* k, size_t n) {
        if (is(k, n, "name")) { str(c, ln.name, ln.name_n); has_name = true; return true; }
        if (is(k, n, "paired_read_name")) { str(c, ln.pair, ln.pair_n); return true; }
        if (is(k, n, "sequence")) { str(c, ln.seq, ln.seq_n); ln.has_seq = true; return true; }
        if (is(k, n, "start")) {
            ln.has_start = true;
            array(c, [&] { ln.start.push_back((int)int_like(c)); });
            return true;
        }
        if (is(k, n, "subpath")) {
            ln.has_subpath = true;
            array(c, [&] {
                Sub sp;
                sp.m0 = sp.m1 = (uint32_t)ln.maps.size();
                sp.n0 = sp.n1 = (uint32_t)ln.nexts.size();
                sp.score = 0.0;
                sp.has_next = false;
                bool has_path = false, has_mapping = false;
                object(c, [&](const char* k2, size_t n2) {
                    if (is(k2, n2, "score")) { ws(c); sp.score = num(c); return true; }
                    if (is(k2, n2, "next")) {
                        sp.has_next = true;
                        array(c, [&] { ln.nexts.push_back((int)int_like(c)); });
                        sp.n1 = (uint32_t)ln.nexts.size();
                        return true;
                    }
                    if (is(k2, n2, "path")) {
                        has_path = true;
                        object(c, [&](const char* k3, size_t n3) {
                            if (!is(k3, n3, "mapping")) return false;
                            has_mapping = true;
                            array(c, [&] {
                                Mapping mp;
                                mp.node = -1;
                                mp.has_edit = false;
                                mp.e0 = mp.e1 = (uint32_t)ln.edits.size();
                                bool has_pos = false, has_node = false;
                                object(c, [&](const char* k4, size_t n4) {
                                    if (is(k4, n4, "position")) {
                                        has_pos = true;
                                        object(c, [&](const char* k5, size_t n5) {
                                            if (!is(k5, n5, "node_id")) return false;
                                            mp.node = int_like(c);
                                            has_node = true;
                                            return true;
                                        });
                                        return true;
                                    }
                                    if (is(k4, n4, "edit")) {
                                        mp.has_edit = true;
                                        array(c, [&] {
                                            Edit ed{0, 0, false};
                                            object(c, [&](const char* k5, size_t n5) {
                                                if (is(k5, n5, "from_length")) { ed.from = (int)int_like(c); return true; }
                                                if (is(k5, n5, "to_length")) { ed.to = (int)int_like(c); return true; }
                                                if (is(k5, n5, "sequence")) { ed.has_seq = true; return false; }
                                                return false;
                                            });
                                            ln.edits.push_back(ed);
                                        });
                                        mp.e1 = (uint32_t)ln.edits.size();
                                        return true;
                                    }
                                    return false;
                                });
                                if (!has_pos || !has_node) mp.node = -2;     // KeyError 'position' / 'node_id' when used
                                ln.maps.push_back(mp);
                            });
                            return true;
                        });
                        return true;
                    }
                    return false;
                });
                sp.m1 = (uint32_t)ln.maps.size();
                if (!has_path || !has_mapping) sp.m0 = sp.m1 = UINT32_MAX;   // KeyError when used
                ln.subs.push_back(sp);
            });
            return true;
        }
        return false;
    });
    if (!has_name) throw Fail{E_KEY, "'name'"};
}


// ---- one GAMP record: vg's MultipathAlignment message (protobuf wire format) ----------------------------------
// Field numbers of vg.proto as published with vg 1.2x (the schema is not vendored in the reference; what it reads of
// the JSON rendering is listed in SURVEY.md section 8c): MultipathAlignment { sequence = 1, name = 3, subpath = 6,
// start = 8, paired_read_name = 9 }, Subpath { path = 1, next = 2, score = 3 }, Path { mapping = 2 },
// Mapping { position = 1, edit = 2 }, Position { node_id = 1 }, Edit { from_length = 1, to_length = 2, sequence = 3 }.
// The same Parsed as a JSON line, with the same presence semantics (`vg view -j` omits fields at their default).
struct Pb {
    const uint8_t* p;
    const uint8_t* e;
    bool more() const { return p < e; }
    uint64_t varint() {
        uint64_t v = 0;
        for (int shift = 0; shift < 70; shift += 7) {
            if (p >= e) throw Fail{E_JSON, "truncated GAMP record"};
            const uint8_t b = *p++;
            v |= (uint64_t)(b & 0x7f) << (shift & 63);
            if (!(b & 0x80)) return v;
        }
        throw Fail{E_JSON, "bad varint in GAMP record"};
    }
    Pb sub() {                                   // a length-delimited field
        const uint64_t n = varint();
        if (n > (uint64_t)(e - p)) throw Fail{E_JSON, "truncated GAMP record"};
        Pb s{p, p + n};
        p += n;
        return s;
    }
    void skip(int wire) {
        switch (wire) {
            case 0: varint(); break;
            case 1: if (e - p < 8) throw Fail{E_JSON, "truncated GAMP record"}; p += 8; break;
            case 2: sub(); break;
            case 5: if (e - p < 4) throw Fail{E_JSON, "truncated GAMP record"}; p += 4; break;
            default: throw Fail{E_JSON, "unsupported wire type in GAMP record"};
        }
    }
};
template <class F>
inline void pb_ints(Pb& c, int wire, F f) {       // a repeated integer field: packed or one by one
    if (wire == 2) { Pb s = c.sub(); while (s.more()) f((long long)s.varint()); }
    else if (wire == 0) f((long long)c.varint());
    else c.skip(wire);
}
void parse_pb(const char* b, const char* e, Parsed& ln) {
    Pb c{(const uint8_t*)b, (const uint8_t*)e};
    ln.reset();
    while (c.more()) {
        const uint64_t key = c.varint();
        const int field = (int)(key >> 3), wire = (int)(key & 7);
        if (wire == 2 && (field == 1 || field == 3 || field == 9)) {
            Pb v = c.sub();
            const char* q = (const char*)v.p;
            const size_t n = (size_t)(v.e - v.p);
            if (field == 1) { ln.seq = q; ln.seq_n = n; ln.has_seq = n > 0; }
            else if (field == 3) { ln.name = q; ln.name_n = n; }
            else { ln.pair = n ? q : nullptr; ln.pair_n = n; }
        } else if (field == 8) {
            ln.has_start = true;
            pb_ints(c, wire, [&](long long v) { ln.start.push_back((int)v); });
        } else if (field == 6 && wire == 2) {
            ln.has_subpath = true;
            Pb sc = c.sub();
            Sub sp;
            sp.m0 = sp.m1 = (uint32_t)ln.maps.size();
            sp.n0 = sp.n1 = (uint32_t)ln.nexts.size();
            sp.score = 0.0;
            sp.has_next = false;
            bool has_path = false, has_mapping = false;
            while (sc.more()) {
                const uint64_t k2 = sc.varint();
                const int f2 = (int)(k2 >> 3), w2 = (int)(k2 & 7);
                if (f2 == 3 && w2 == 0) {
                    sp.score = (double)(int32_t)sc.varint();              // int32: negative values are sign-extended varints
                } else if (f2 == 2) {
                    pb_ints(sc, w2, [&](long long v) { ln.nexts.push_back((int)v); sp.has_next = true; });
                    sp.n1 = (uint32_t)ln.nexts.size();
                } else if (f2 == 1 && w2 == 2) {
                    has_path = true;
                    Pb pc = sc.sub();
                    while (pc.more()) {
                        const uint64_t k3 = pc.varint();
                        if ((k3 >> 3) != 2 || (k3 & 7) != 2) { pc.skip((int)(k3 & 7)); continue; }
                        has_mapping = true;
                        Pb mc = pc.sub();
                        Mapping mp;
                        mp.node = -1;
                        mp.has_edit = false;
                        mp.e0 = mp.e1 = (uint32_t)ln.edits.size();
                        bool has_pos = false, has_node = false;
                        while (mc.more()) {
                            const uint64_t k4 = mc.varint();
                            const int f4 = (int)(k4 >> 3), w4 = (int)(k4 & 7);
                            if (f4 == 1 && w4 == 2) {
                                has_pos = true;
                                Pb qc = mc.sub();
                                while (qc.more()) {
                                    const uint64_t k5 = qc.varint();
                                    if ((k5 >> 3) == 1 && (k5 & 7) == 0) { mp.node = (long long)qc.varint(); has_node = true; }
                                    else qc.skip((int)(k5 & 7));
                                }
                            } else if (f4 == 2 && w4 == 2) {
                                mp.has_edit = true;
                                Pb ec = mc.sub();
                                Edit ed{0, 0, false};
                                while (ec.more()) {
                                    const uint64_t k5 = ec.varint();
                                    const int f5 = (int)(k5 >> 3), w5 = (int)(k5 & 7);
                                    if (f5 == 1 && w5 == 0) ed.from = (int)ec.varint();
                                    else if (f5 == 2 && w5 == 0) ed.to = (int)ec.varint();
                                    else if (f5 == 3 && w5 == 2) { Pb t = ec.sub(); ed.has_seq = t.e > t.p; }
                                    else ec.skip(w5);
                                }
                                ln.edits.push_back(ed);
                                mp.e1 = (uint32_t)ln.edits.size();
                            } else {
                                mc.skip(w4);
                            }
                        }
                        if (!has_pos || !has_node) mp.node = -2;          // KeyError 'position' / 'node_id' when used
                        ln.maps.push_back(mp);
                    }
                } else {
                    sc.skip(w2);
                }
            }
            sp.m1 = (uint32_t)ln.maps.size();
            if (!has_path || !has_mapping) sp.m0 = sp.m1 = UINT32_MAX;        // KeyError when used
            ln.subs.push_back(sp);
        } else {
            c.skip(wire);
        }
    }
    if (!ln.name) throw Fail{E_KEY, "'name'"};           // `vg view -j` prints no "name" key for an empty name
}

struct NodeTable {
    const int64_t* ids;
    const int32_t* len;
    int64_t n;
    struct Slot { int32_t dense1, len; };        // dense index + 1 (0: no such node) and the node's length, one cache line access
    std::vector<Slot> direct;                    // by id when ids are small
    std::unordered_map<int64_t, int32_t> sparse;
    void build() {
        int64_t mx = -1;
        bool neg = false;
        for (int64_t i = 0; i < n; ++i) { mx = std::max(mx, ids[i]); neg |= ids[i] < 0; }
        if (!neg && mx < 8 * n + 1024) {
            direct.assign((size_t)mx + 1, Slot{0, 0});
            for (int64_t i = 0; i < n; ++i) direct[(size_t)ids[i]] = Slot{(int32_t)i + 1, len[i]};
        } else {
            sparse.reserve((size_t)n * 2);
            for (int64_t i = 0; i < n; ++i) sparse[ids[i]] = (int32_t)i;
        }
    }
    int32_t find(long long id, int32_t& node_len) const {
        if (!direct.empty()) {
            if (id < 0 || (size_t)id >= direct.size() || direct[(size_t)id].dense1 == 0) throw Fail{E_KEY, std::to_string(id)};
            node_len = direct[(size_t)id].len;
            return direct[(size_t)id].dense1 - 1;
        }
        auto it = sparse.find(id);
        if (it == sparse.end()) throw Fail{E_KEY, std::to_string(id)};
        node_len = len[it->second];
        return it->second;
    }
};

inline long long sub_to_len(const Parsed& ln, int s) {
    if (s < 0 || (size_t)s >= ln.subs.size()) throw Fail{E_INDEX, "list index out of range"};
    const Sub& sp = ln.subs[s];
    if (sp.m0 == UINT32_MAX) throw Fail{E_KEY, "'path'"};
    long long t = 0;
    for (uint32_t m = sp.m0; m < sp.m1; ++m) {
        if (!ln.maps[m].has_edit) throw Fail{E_KEY, "'edit'"};
        for (uint32_t e = ln.maps[m].e0; e < ln.maps[m].e1; ++e) t += ln.edits[e].to;
    }
    return t;
}

// BFS (json2csv.py:589-662): FIFO relaxation keeping, per end metanode, the best-scoring routes.  The
// reference's dict of live routes is insertion ordered: `order` (entries in insertion order, dead ones
// flagged) + `slot` (metanode -> live entry) reproduce that ordering.  All storage is per thread and reused.
struct Route {
    uint32_t p0, pn;       // metanodes of the route: range in Router::paths
    long long read_len;
    double score;
};
struct Router {
    struct Entry {
        int key;
        bool alive;
        std::vector<uint32_t> routes;    // indices into `routes`
    };
    std::vector<Entry> order;
    size_t n_order = 0;
    std::vector<int> slot, fifo;
    std::vector<Route> routes;
    std::vector<int> paths;
    std::vector<uint32_t> best;          // result: indices into `routes`

    int add_entry(int key) {
        if (n_order == order.size()) order.emplace_back();
        Entry& en = order[n_order];
        en.key = key;
        en.alive = true;
        en.routes.clear();
        return (int)n_order++;
    }
    uint32_t extend(uint32_t parent, int child, long long c_len, double c_score) {
        const Route pr = routes[parent];
        const uint32_t p0 = (uint32_t)paths.size();
        paths.resize(p0 + pr.pn + 1);
        std::copy(paths.begin() + pr.p0, paths.begin() + pr.p0 + pr.pn, paths.begin() + p0);
        paths[p0 + pr.pn] = child;
        routes.push_back(Route{p0, pr.pn + 1, pr.read_len + c_len, pr.score + c_score});
        return (uint32_t)routes.size() - 1;
    }

    void run(const Parsed& ln) {
        if (!ln.has_start) throw Fail{E_KEY, "'start'"};
        if (!ln.has_seq) throw Fail{E_KEY, "'sequence'"};
        n_order = 0;
        routes.clear();
        paths.clear();
        best.clear();
        fifo.assign(ln.start.begin(), ln.start.end());
        slot.assign(ln.subs.size(), -1);
        size_t head = 0;
        for (int s : ln.start) {
            const long long len = sub_to_len(ln, s);            // also the range check of s
            paths.push_back(s);
            routes.push_back(Route{(uint32_t)paths.size() - 1, 1, len, ln.subs[s].score});
            if (slot[s] < 0) slot[s] = add_entry(s);
            order[slot[s]].routes.assign(1, (uint32_t)routes.size() - 1);
        }
        const long long want = (long long)ln.seq_n;
        while (head < fifo.size()) {
            const int cur = fifo[head++];
            if (cur < 0 || (size_t)cur >= ln.subs.size()) throw Fail{E_INDEX, "list index out of range"};
            const Sub& sp = ln.subs[cur];
            if (!sp.has_next) continue;
            if (slot[cur] < 0) throw Fail{E_KEY, std::to_string(cur)};
            bool all_done = true;
            for (uint32_t r : order[slot[cur]].routes) all_done &= routes[r].read_len == want;
            if (all_done) continue;
            for (uint32_t ni = sp.n0; ni < sp.n1; ++ni) {
                const int child = ln.nexts[ni];
                if (std::find(fifo.begin() + head, fifo.end(), child) == fifo.end()) fifo.push_back(child);
                const long long c_len = sub_to_len(ln, child);
                const double c_score = ln.subs[child].score;
                if (slot[child] < 0) slot[child] = add_entry(child);
                const int ci = slot[child], pi = slot[cur];
                const size_t n_parent = order[pi].routes.size();      // child == cur would grow it while iterating
                for (size_t q = 0; q < n_parent && q < order[pi].routes.size(); ++q) {
                    const uint32_t nr = extend(order[pi].routes[q], child, c_len, c_score);
                    std::vector<uint32_t>& bucket = order[ci].routes;
                    if (bucket.empty()) bucket.push_back(nr);
                    else if (routes[nr].score > routes[bucket[0]].score) bucket.assign(1, nr);
                    else if (routes[nr].score == routes[bucket[0]].score) bucket.push_back(nr);
                }
            }
            Entry& done = order[slot[cur]];
            done.alive = false;
            done.routes.clear();
            slot[cur] = -1;
        }
        size_t alive = 0;
        double top = 0.0;
        bool first = true;
        for (size_t i = 0; i < n_order; ++i) {
            const Entry& en = order[i];
            if (!en.alive) continue;
            ++alive;
            if (en.routes.empty()) throw Fail{E_INDEX, "list index out of range"};
            const double sc = routes[en.routes[0]].score;
            if (first || sc > top) top = sc;
            first = false;
        }
        for (size_t i = 0; i < n_order; ++i) {
            const Entry& en = order[i];
            if (!en.alive) continue;
            if (alive > 1 && routes[en.routes[0]].score < top) continue;
            for (uint32_t r : en.routes) best.push_back(r);
        }
    }
};

// Python round(x, 2) -> histogram key index, 255 when outside the table (json2csv.py:269-273, 335-339)
inline uint8_t bin_of(double x) {
    // away from a rounding tie the nearest multiple of 0.01 is unambiguous; x * 100 is exact to ~1e-14 here
    const double y = x * 100.0, f = std::floor(y), frac = y - f;
    if (std::fabs(frac - 0.5) > 1e-6 && y > -1e6 && y < 1e6) {
        const long k = (long)(frac > 0.5 ? f + 1.0 : f);
        return (k < 0 || k > 100) ? (uint8_t)255 : (uint8_t)k;
    }
    char buf[64];
    snprintf(buf, sizeof buf, "%.2f", x);          // glibc: exact decimal expansion, round-half-even (= Python round)
    const double r = strtod(buf, nullptr);
    const long k = lround(r * 100.0);
    if (k < 0 || k > 100 || (double)k / 100.0 != r) return 255;
    return (uint8_t)k;
}

// ---- per-thread results of phase A -------------------------------------------------------------------
struct AlnOut {          // one decoded Alignment (json2csv.py:122-134); its records are [rec0, rec0 + nrec) of the arena
    uint32_t rec0, nrec;
    int32_t read_len;
    double stats[5];
    uint8_t bins[5];
};
struct Arena {
    std::vector<AlnOut> alns;
    std::vector<int32_t> node, alen;     // dense node index, aligned bases
    std::vector<double> cov;
    std::vector<Fail> fails;
    std::vector<char> text;              // read names (raw JSON bytes), packed: phase B compares them without touching the file again
    void clear() { alns.clear(); node.clear(); alen.clear(); cov.clear(); fails.clear(); text.clear(); }
};
struct LineOut {
    const char* line;        // start of the line in the file
    const char* seq;
    uint32_t seq_n;
    uint32_t name_off, name_n, pair_off, pair_n;     // in arena[tid].text
    uint32_t aln0, naln;     // in arena[tid].alns
    int32_t fail;            // index in arena[tid].fails, -1: none
    uint16_t tid;
    bool has_name, has_seq, has_subpath;
};
struct Worker {
    Parsed parsed;
    Router router;
    Arena arena;
    std::vector<uint32_t> sel, s0, s1;   // phase C: lines whose alignments the mates keep
    std::string text;
};

// metaalign2align (json2csv.py:551-587), appended to the arena
void route_to_aln(const Parsed& ln, const Router& rt, const Route& route, const NodeTable& nt, Arena& ar) {
    AlnOut a;
    a.rec0 = (uint32_t)ar.node.size();
    long long tot_al = 0, tot_match = 0, tot_indel = 0;
    for (uint32_t pi = route.p0; pi < route.p0 + route.pn; ++pi) {
        const Sub& sp = ln.subs[rt.paths[pi]];
        for (uint32_t m = sp.m0; m < sp.m1; ++m) {
            const Mapping& mp = ln.maps[m];
            if (mp.node == -2) throw Fail{E_KEY, "'position'"};
            int32_t nl = 0;
            const int32_t dense = nt.find(mp.node, nl);
            if (!mp.has_edit) throw Fail{E_KEY, "'edit'"};
            long long n_al = 0, n_match = 0;
            for (uint32_t e = mp.e0; e < mp.e1; ++e) {
                const Edit& ed = ln.edits[e];
                tot_indel += std::llabs((long long)ed.from - ed.to);
                n_al += std::min(ed.from, ed.to);
                if (ed.from == ed.to && !ed.has_seq) n_match += ed.to;
            }
            tot_al += n_al;
            tot_match += n_match;
            if (nl == 0) throw Fail{E_ZERODIV, "division by zero"};
            const double frac = (double)n_al / (double)nl;
            if (ar.node.size() > a.rec0 && ar.node.back() == dense && ar.cov.back() < 1.0) {
                ar.cov.back() += frac;
                ar.alen.back() += (int32_t)n_al;
            } else {
                ar.node.push_back(dense);
                ar.cov.push_back(frac);
                ar.alen.push_back((int32_t)n_al);
            }
        }
    }
    const double rl = (double)route.read_len;
    if (route.read_len == 0 || tot_al == 0) throw Fail{E_ZERODIV, "division by zero"};
    const long long tot_subst = tot_al - tot_match;
    a.nrec = (uint32_t)ar.node.size() - a.rec0;
    a.read_len = (int32_t)route.read_len;
    a.stats[0] = route.score / rl;
    a.stats[1] = (double)tot_match / (double)tot_al;
    a.stats[2] = (double)tot_subst / rl;
    a.stats[3] = (double)tot_indel / rl;
    a.stats[4] = (double)(tot_indel + tot_subst) / rl;
    for (int k = 0; k < 5; ++k) a.bins[k] = bin_of(a.stats[k]);
    ar.alns.push_back(a);
}

void decode_line(const char* b, const char* e, Worker& w, int tid, const NodeTable& nt, LineOut& out, bool pb = false) {
    Arena& ar = w.arena;
    out.line = b;
    out.fail = -1;
    out.tid = (uint16_t)tid;
    out.aln0 = (uint32_t)ar.alns.size();
    out.naln = 0;
    const size_t rec_mark = ar.node.size();
    try {
        if (pb) parse_pb(b, e, w.parsed); else parse_line(b, e, w.parsed);
        if (w.parsed.has_subpath) {
            w.router.run(w.parsed);
            for (uint32_t r : w.router.best) route_to_aln(w.parsed, w.router, w.router.routes[r], nt, ar);
            if (w.router.best.empty()) throw Fail{E_INDEX, "list index out of range"};
            out.naln = (uint32_t)ar.alns.size() - out.aln0;
        }
    } catch (const Fail& f) {
        ar.alns.resize(out.aln0);
        ar.node.resize(rec_mark);
        ar.alen.resize(rec_mark);
        ar.cov.resize(rec_mark);
        out.naln = 0;
        out.fail = (int32_t)ar.fails.size();
        ar.fails.push_back(f);
    }
    const Parsed& p = w.parsed;
    out.seq = p.seq;
    out.seq_n = (uint32_t)p.seq_n;
    out.has_name = p.name != nullptr;
    // names are kept as JSON text (escapes are undone when they are written out); raw protobuf names are brought to
    // that form by doubling their backslashes
    auto put = [&](const char* q, size_t n, uint32_t& off, uint32_t& len) {
        off = (uint32_t)ar.text.size();
        if (!pb || !n || !memchr(q, '\\', n)) {
            if (n) ar.text.insert(ar.text.end(), q, q + n);
        } else {
            for (size_t i = 0; i < n; ++i) {
                if (q[i] == '\\') ar.text.push_back('\\');
                ar.text.push_back(q[i]);
            }
        }
        len = (uint32_t)(ar.text.size() - off);
    };
    put(p.name, p.name_n, out.name_off, out.name_n);
    put(p.pair, p.pair_n, out.pair_off, out.pair_n);
    out.has_seq = p.has_seq;
    out.has_subpath = p.has_subpath;
}

// ---- output --------------------------------------------------------------------------------------------
template <class T>
struct Buf {             // exact-size, uninitialised storage
    T* p = nullptr;
    size_t n = 0;
    Buf() = default;
    Buf(const Buf&) = delete;
    Buf& operator=(const Buf&) = delete;
    Buf(Buf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    Buf& operator=(Buf&& o) noexcept { std::swap(p, o.p); std::swap(n, o.n); return *this; }
    ~Buf() { free(p); }
    void alloc(size_t count) {
        free(p);
        n = count;
        p = count ? (T*)malloc(count * sizeof(T)) : nullptr;
        if (count && !p) throw Fail{E_IO, "out of memory"};
    }
};
struct OutBlock {        // the decoded read groups of one block of the file; offsets are global
    int64_t G = 0, A = 0, R = 0, NB = 0;
    Buf<uint8_t> grp_nmates, aln_bins;
    Buf<int64_t> grp_aln_off, aln_walk_off, name_off;     // [2G], [A], [2G]: the END offsets (the leading 0 is implicit)
    Buf<int32_t> mate_seq_len, aln_read_len, walk_node, walk_alen;
    Buf<double> aln_stats;
    Buf<char> names;
};

}  // namespace

struct sfb_decoded {
    std::vector<OutBlock> blocks;
    std::vector<double> exc_cov;
    int err_kind = E_NONE;
    std::string err_msg;
    int64_t n_groups = 0, n_alns = 0, n_recs = 0, n_name_bytes = 0;
};

namespace {

inline bool same(const char* a, size_t an, const char* b, size_t bn) { return an == bn && memcmp(a, b, an) == 0; }

void unescape_into(std::string& out, const char* b, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        if (b[i] != '\\' || i + 1 >= n) { out.push_back(b[i]); continue; }
        const char c = b[++i];
        switch (c) {
            case 'n': out.push_back('\n'); break;
            case 't': out.push_back('\t'); break;
            case 'r': out.push_back('\r'); break;
            case 'b': out.push_back('\b'); break;
            case 'f': out.push_back('\f'); break;
            case 'u': {
                unsigned cp = 0;
                if (i + 4 < n + 0 && sscanf(std::string(b + i + 1, 4).c_str(), "%4x", &cp) == 1) i += 4;
                if (cp < 0x80) out.push_back((char)cp);
                else if (cp < 0x800) { out.push_back((char)(0xC0 | (cp >> 6))); out.push_back((char)(0x80 | (cp & 0x3F))); }
                else { out.push_back((char)(0xE0 | (cp >> 12))); out.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out.push_back((char)(0x80 | (cp & 0x3F))); }
                break;
            }
            default: out.push_back(c);
        }
    }
}
// length of the unescaped name / the name itself (names without a backslash, i.e. all real ones, are copied as they are)
inline size_t name_len(const char* b, size_t n, std::string& tmp) {
    if (!n || !memchr(b, '\\', n)) return n;
    tmp.clear();
    unescape_into(tmp, b, n);
    return tmp.size();
}
inline void name_put(const char* b, size_t n, std::string& tmp, char* dst) {
    if (!n) return;
    if (!memchr(b, '\\', n)) { memcpy(dst, b, n); return; }
    tmp.clear();
    unescape_into(tmp, b, n);
    memcpy(dst, tmp.data(), tmp.size());
}

struct Group {           // a read group: lines [l0, l1) of the block
    uint32_t l0, l1;
    uint32_t mate1_line;             // the line that opened the second mate
    uint32_t sel_off, n_sel[2];      // kept lines per mate: worker[tid].sel[sel_off ...], mate 0 first
    uint32_t n_aln[2], n_rec, name_bytes[2];
    uint16_t tid;
    uint8_t nm;
    int64_t aln_base, rec_base, name_base;     // global offsets (phase D)
};

constexpr size_t kPiece = 256u << 10;      // bytes of the file one work item of phase A covers (at most)

struct Exc {             // a record whose coverage has to be shipped as fp64
    int64_t rec;
    double cov;
};

// Phase B (json2csv.py:667-739, the loop structure): read-group boundaries.  Throws what the reference would
// raise first.  Returns the index of the first line NOT consumed: when `eof` is false the last group of the
// block may continue in the next block.
size_t find_groups(const std::vector<LineOut>& lines, const std::vector<Worker>& workers, bool eof, std::vector<Group>& groups) {
    const size_t n = lines.size();
    auto fail_of = [&](const LineOut& ln) -> const Fail& { return workers[ln.tid].arena.fails[(size_t)ln.fail]; };
    auto text_of = [&](const LineOut& ln, uint32_t off) { return workers[ln.tid].arena.text.data() + off; };
    size_t i = 0;
    while (i < n) {
        const LineOut& first = lines[i];
        if (first.fail >= 0) throw fail_of(first);
        if (!first.has_seq) throw Fail{E_KEY, "'sequence'"};
        const char* f_name = text_of(first, first.name_off);
        const char* f_pair = text_of(first, first.pair_off);
        size_t j = i + 1;
        for (; j < n; ++j) {
            const LineOut& ln = lines[j];
            if (ln.fail >= 0 && fail_of(ln).kind == E_JSON) throw fail_of(ln);      // json.loads comes before the name test
            if (!ln.has_name && ln.fail < 0) throw Fail{E_KEY, "'name'"};
            const char* l_name = text_of(ln, ln.name_off);
            if (!same(l_name, ln.name_n, f_name, first.name_n) && !same(l_name, ln.name_n, f_pair, first.pair_n)) break;
            if (ln.fail >= 0) throw fail_of(ln);
            if (!ln.has_subpath) continue;
            if (!ln.has_seq) throw Fail{E_KEY, "'sequence'"};
        }
        if (j == n && !eof) return i;
        Group g{};
        g.l0 = (uint32_t)i;
        g.l1 = (uint32_t)j;
        groups.push_back(g);
        i = j;
    }
    return n;
}

// Phase C (json2csv.py:684-737): the alignments a mate keeps are those of the lines with the best score seen,
// provided the first line seen for the mate passes the threshold ... exactly the reference's update rule.
void select_group(Group& g, const std::vector<LineOut>& lines, std::vector<Worker>& workers, int tid, double thr) {
    Worker& w = workers[tid];
    auto score_of = [&](uint32_t li) { const LineOut& ln = lines[li]; return workers[ln.tid].arena.alns[ln.aln0].stats[0]; };
    const LineOut& first = lines[g.l0];
    std::vector<uint32_t>* sel[2] = {&w.s0, &w.s1};
    w.s0.clear();
    w.s1.clear();
    if (first.has_subpath && score_of(g.l0) >= thr) w.s0.push_back(g.l0);
    g.nm = 1;
    g.mate1_line = g.l0;
    for (uint32_t j = g.l0 + 1; j < g.l1; ++j) {
        const LineOut& ln = lines[j];
        if (!ln.has_subpath) continue;
        const bool is_first_seq = same(ln.seq, ln.seq_n, first.seq, first.seq_n);
        if (!is_first_seq && g.nm == 1) { g.nm = 2; g.mate1_line = j; }
        std::vector<uint32_t>& tgt = *sel[is_first_seq ? 0 : 1];
        const double best = score_of(j);
        if (tgt.empty()) {
            if (best >= thr) tgt.push_back(j);
        } else {
            const double have = score_of(tgt[0]);
            if (have < best) tgt.assign(1, j);
            else if (have == best) tgt.push_back(j);
        }
    }
    g.tid = (uint16_t)tid;
    g.sel_off = (uint32_t)w.sel.size();
    g.n_rec = 0;
    for (int m = 0; m < 2; ++m) {
        g.n_sel[m] = (uint32_t)sel[m]->size();
        g.n_aln[m] = 0;
        for (uint32_t li : *sel[m]) {
            const LineOut& ln = lines[li];
            w.sel.push_back(li);
            g.n_aln[m] += ln.naln;
            const Arena& ar = workers[ln.tid].arena;
            for (uint32_t a = 0; a < ln.naln; ++a) g.n_rec += ar.alns[ln.aln0 + a].nrec;
        }
    }
    auto raw_name = [&](const LineOut& ln) { return workers[ln.tid].arena.text.data() + ln.name_off; };
    g.name_bytes[0] = (uint32_t)name_len(raw_name(first), first.name_n, w.text);
    const LineOut& second = lines[g.mate1_line];
    g.name_bytes[1] = g.nm == 2 ? (uint32_t)name_len(raw_name(second), second.name_n, w.text) : 0u;
}

// Phase E: the group's part of the output arrays.  `gi` is the group's index inside the block; b0/r0/n0 the block's bases.
void fill_group(const Group& g, size_t gi, OutBlock& ob, int64_t a0, int64_t r0, int64_t n0, const std::vector<LineOut>& lines,
                std::vector<Worker>& workers, int tid, const int32_t* node_len, std::vector<Exc>& exc) {
    Worker& me = workers[tid];
    const uint32_t* sel = workers[g.tid].sel.data() + g.sel_off;
    ob.grp_nmates.p[gi] = g.nm;
    int64_t aln = g.aln_base, rec = g.rec_base, nb = g.name_base;
    for (int m = 0; m < 2; ++m) {
        if (m < (int)g.nm) {
            const LineOut& src = lines[m == 0 ? g.l0 : g.mate1_line];
            ob.mate_seq_len.p[2 * gi + m] = (int32_t)src.seq_n;
            name_put(workers[src.tid].arena.text.data() + src.name_off, src.name_n, me.text, ob.names.p + (nb - n0));
            nb += g.name_bytes[m];
            for (uint32_t s = 0; s < g.n_sel[m]; ++s) {
                const LineOut& ln = lines[*sel++];
                const Arena& ar = workers[ln.tid].arena;
                for (uint32_t k = 0; k < ln.naln; ++k) {
                    const AlnOut& a = ar.alns[ln.aln0 + k];
                    const int64_t la = aln - a0;
                    ob.aln_read_len.p[la] = a.read_len;
                    for (int q = 0; q < 5; ++q) {
                        ob.aln_stats.p[la * 5 + q] = a.stats[q];
                        ob.aln_bins.p[la * 5 + q] = a.bins[q];
                    }
                    for (uint32_t i = 0; i < a.nrec; ++i) {
                        const int64_t lr = rec - r0 + i;
                        const int32_t nd = ar.node[a.rec0 + i], al = ar.alen[a.rec0 + i], nl = node_len[nd];
                        const double cov = ar.cov[a.rec0 + i];
                        ob.walk_node.p[lr] = nd;
                        // aligned bases | node length << 16 when that reproduces the reference's fp64, else an explicit entry
                        if (al >= 0 && al <= 0xFFFF && nl >= 1 && nl <= 0x7FFF && (double)al / (double)nl == cov) {
                            ob.walk_alen.p[lr] = al | (nl << 16);
                        } else {
                            ob.walk_alen.p[lr] = 0;
                            exc.push_back(Exc{rec + i, cov});
                        }
                    }
                    rec += a.nrec;
                    ob.aln_walk_off.p[la] = rec;
                    aln += 1;
                }
            }
        } else {
            ob.mate_seq_len.p[2 * gi + m] = 0;
        }
        ob.grp_aln_off.p[2 * gi + m] = aln;
        ob.name_off.p[2 * gi + m] = nb;
    }
}

template <class F>
void run_parallel(int T, size_t n_items, size_t grain, F&& body) {
    const int use = (int)std::max<size_t>(1, std::min<size_t>((size_t)T, (n_items + grain - 1) / grain));
    std::atomic<size_t> next{0};
    auto work = [&](int tid) {
        for (;;) {
            const size_t lo = next.fetch_add(grain);
            if (lo >= n_items) return;
            body(lo, std::min(n_items, lo + grain), tid);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < use; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
}

// Phases B-E for the decoded records of one block (JSON lines or GAMP messages): read-group boundaries, what every mate
// keeps, offsets, fill.  Returns the number of records consumed; 0 (with eof false): one read group larger than the block.
struct BlockTimes { double b, c, e; };
size_t finish_block(sfb_decoded* D, std::vector<LineOut>& lines, std::vector<Group>& groups, std::vector<Worker>& workers,
                    std::vector<std::vector<Exc>>& exc, int T, double thr, const int32_t* node_len, bool eof, BlockTimes& bt) {
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t0;
    double &t_b = bt.b, &t_c = bt.c, &t_e = bt.e;
    // B: groups
    t0 = now();
    groups.clear();
    const size_t used = find_groups(lines, workers, eof, groups);
    t_b += now() - t0;
    if (used == 0 && !eof) return 0;       // one read group larger than the block: the caller widens the block
    // C: what every mate keeps
    t0 = now();
    run_parallel(T, groups.size(), 512, [&](size_t lo, size_t hi, int tid) {
        for (size_t g = lo; g < hi; ++g) select_group(groups[g], lines, workers, tid, thr);
    });
    // D: offsets
    OutBlock ob;
    ob.G = (int64_t)groups.size();
    int64_t aln = D->n_alns, rec = D->n_recs, nb = D->n_name_bytes;
    for (Group& g : groups) {
        g.aln_base = aln;
        g.rec_base = rec;
        g.name_base = nb;
        aln += g.n_aln[0] + g.n_aln[1];
        rec += g.n_rec;
        nb += g.name_bytes[0] + g.name_bytes[1];
    }
    ob.A = aln - D->n_alns;
    ob.R = rec - D->n_recs;
    ob.NB = nb - D->n_name_bytes;
    t_c += now() - t0;
    // E: fill
    t0 = now();
    ob.grp_nmates.alloc((size_t)ob.G);
    ob.grp_aln_off.alloc((size_t)ob.G * 2);
    ob.name_off.alloc((size_t)ob.G * 2);
    ob.mate_seq_len.alloc((size_t)ob.G * 2);
    ob.aln_walk_off.alloc((size_t)ob.A);
    ob.aln_read_len.alloc((size_t)ob.A);
    ob.aln_bins.alloc((size_t)ob.A * 5);
    ob.aln_stats.alloc((size_t)ob.A * 5);
    ob.walk_node.alloc((size_t)ob.R);
    ob.walk_alen.alloc((size_t)ob.R);
    ob.names.alloc((size_t)ob.NB);
    const int64_t a0 = D->n_alns, r0 = D->n_recs, n0 = D->n_name_bytes;
    run_parallel(T, groups.size(), 512, [&](size_t lo, size_t hi, int tid) {
        for (size_t g = lo; g < hi; ++g)
            fill_group(groups[g], g, ob, a0, r0, n0, lines, workers, tid, node_len, exc[(size_t)tid]);
    });
    // explicit coverages in record order (deterministic whatever the thread count)
    std::vector<Exc> all;
    for (auto& v : exc) { all.insert(all.end(), v.begin(), v.end()); v.clear(); }
    std::sort(all.begin(), all.end(), [](const Exc& x, const Exc& y) { return x.rec < y.rec; });
    for (const Exc& e : all) {
        D->exc_cov.push_back(e.cov);
        ob.walk_alen.p[(size_t)(e.rec - r0)] = -(int32_t)D->exc_cov.size();
    }
    D->n_groups += ob.G;
    D->n_alns = aln;
    D->n_recs = rec;
    D->n_name_bytes = nb;
    D->blocks.push_back(std::move(ob));
    t_e += now() - t0;
    return used;
}

}  // namespace

extern "C" {

sfb_decoded* sfb_decode_json(const char* path, const int64_t* node_ids, const int32_t* node_len, int64_t n_nodes,
                             double thr, int n_threads, int64_t block_bytes) {
    sfb_decoded* D = new sfb_decoded();
    int fd = -1;
    const char* data = nullptr;
    size_t size = 0;
    try {
        fd = open(path, O_RDONLY);
        if (fd < 0) throw Fail{E_IO, std::string("cannot open ") + path};
        struct stat sb;
        if (fstat(fd, &sb) != 0) throw Fail{E_IO, "fstat failed"};
        size = (size_t)sb.st_size;
        if (size) {
            void* m = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
            if (m == MAP_FAILED) throw Fail{E_IO, "mmap failed"};
            data = (const char*)m;
            madvise(m, size, MADV_SEQUENTIAL);
        }
        if (block_bytes <= 0) block_bytes = 64ll << 20;
        auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
        const bool trace = getenv("SFB_DECODE_TRACE") != nullptr;
        double t_a = 0, t_b = 0, t_c = 0, t_e = 0, t0;
        NodeTable nt{node_ids, node_len, n_nodes, {}, {}};
        nt.build();
        const int T = std::max(1, std::min(n_threads, 4096));
        std::vector<Worker> workers((size_t)T);
        std::vector<std::vector<Exc>> exc((size_t)T);
        std::vector<size_t> cuts, line_base;     // a block is cut into pieces of whole lines; first line of every piece
        std::vector<LineOut> lines;
        std::vector<Group> groups;
        size_t pos = 0;
        while (pos < size) {
            // one block of whole lines
            size_t end = std::min(size, pos + (size_t)block_bytes);
            if (end < size) {
                size_t cut = end;
                while (cut > pos && data[cut - 1] != '\n') --cut;
                if (cut == pos) {                  // a line longer than the block: take it whole
                    const void* nl = memchr(data + end, '\n', size - end);
                    cut = nl ? (size_t)((const char*)nl - data) + 1 : size;
                }
                end = cut;
            }
            const bool eof = end >= size;
            // A: decode the lines (count them first, so that every piece knows where its lines go)
            t0 = now();
            cuts.clear();
            const size_t piece = std::min(kPiece, std::max<size_t>(1024, (size_t)block_bytes / 64));
            for (size_t p = pos; p < end;) {
                cuts.push_back(p);
                size_t q = std::min(end, p + piece);
                if (q < end) {
                    const void* nl = memchr(data + q, '\n', end - q);
                    q = nl ? (size_t)((const char*)nl - data) + 1 : end;
                }
                p = q;
            }
            const size_t n_pieces = cuts.size();
            cuts.push_back(end);
            line_base.assign(n_pieces + 1, 0);
            run_parallel(T, n_pieces, 1, [&](size_t lo, size_t hi, int) {
                for (size_t c = lo; c < hi; ++c) {
                    size_t count = 0;
                    for (size_t p = cuts[c]; p < cuts[c + 1]; ++count) {
                        const void* nl = memchr(data + p, '\n', cuts[c + 1] - p);
                        p = nl ? (size_t)((const char*)nl - data) + 1 : cuts[c + 1];
                    }
                    line_base[c + 1] = count;
                }
            });
            for (size_t c = 0; c < n_pieces; ++c) line_base[c + 1] += line_base[c];
            const size_t n_lines = line_base[n_pieces];
            for (Worker& w : workers) { w.arena.clear(); w.sel.clear(); }
            lines.resize(n_lines);
            run_parallel(T, n_pieces, 1, [&](size_t lo, size_t hi, int tid) {
                for (size_t c = lo; c < hi; ++c) {
                    size_t i = line_base[c];
                    for (size_t p = cuts[c]; p < cuts[c + 1]; ++i) {
                        const void* nl = memchr(data + p, '\n', cuts[c + 1] - p);
                        const size_t q = nl ? (size_t)((const char*)nl - data) : cuts[c + 1];
                        decode_line(data + p, data + q, workers[(size_t)tid], tid, nt, lines[i]);
                        p = q + 1;
                    }
                }
            });
            t_a += now() - t0;
            BlockTimes bt{0, 0, 0};
            const size_t used = finish_block(D, lines, groups, workers, exc, T, thr, node_len, eof, bt);
            t_b += bt.b; t_c += bt.c; t_e += bt.e;
            if (used == 0 && !eof) {               // one read group larger than the block: widen the block
                block_bytes *= 2;
                continue;
            }
            pos = used < n_lines ? (size_t)(lines[used].line - data) : end;
        }
        if (trace)
            fprintf(stderr, "[decode] A lines (parallel) %.3f s, B groups %.3f s, C+D select (parallel) %.3f s, E fill (parallel) %.3f s, %zu blocks\n",
                    t_a, t_b, t_c, t_e, D->blocks.size());
    } catch (const Fail& f) {
        D->err_kind = f.kind;
        D->err_msg = f.msg;
    } catch (const std::exception& ex) {
        D->err_kind = E_IO;
        D->err_msg = ex.what();
    }
    if (data) munmap((void*)data, size);
    if (fd >= 0) close(fd);
    return D;
}


// ---- GAMP: the file `vg mpmap` writes (StrainFLAIR.sh:582), without the `vg view -j -K` hop (:592-602) ---------------
// Container (vg's stream framing, libvgio): optionally BGZF / gzip compressed; a sequence of groups, each a varint
// message count followed by that many (varint length, bytes) messages; in the type-tagged flavour the first message
// of every group is the tag "MGAM".  Restated from the published format: no vg binary, schema or GAMP file is
// available in the reference checkout or in this image -- parity unpinned (the tests write GAMP files with the
// protobuf library from the golden JSON cases and require decode(GAMP) == decode(JSON)).
sfb_decoded* sfb_decode_gamp(const char* path, const int64_t* node_ids, const int32_t* node_len, int64_t n_nodes,
                             double thr, int n_threads, int64_t block_records) {
    sfb_decoded* D = new sfb_decoded();
    int fd = -1;
    const char* map = nullptr;
    size_t map_size = 0;
    std::vector<char> plain;                      // the decompressed stream (the records point into it)
    try {
        fd = open(path, O_RDONLY);
        if (fd < 0) throw Fail{E_IO, std::string("cannot open ") + path};
        struct stat sb;
        if (fstat(fd, &sb) != 0) throw Fail{E_IO, "fstat failed"};
        map_size = (size_t)sb.st_size;
        if (map_size) {
            void* m = mmap(nullptr, map_size, PROT_READ, MAP_PRIVATE, fd, 0);
            if (m == MAP_FAILED) throw Fail{E_IO, "mmap failed"};
            map = (const char*)m;
        }
        const char* data = map;
        size_t size = map_size;
        if (size >= 2 && (uint8_t)map[0] == 0x1f && (uint8_t)map[1] == 0x8b) {
            // gzip members one after the other (BGZF is that, with 64 KiB blocks)
            z_stream zs;
            memset(&zs, 0, sizeof zs);
            if (inflateInit2(&zs, 16 + MAX_WBITS) != Z_OK) throw Fail{E_IO, "zlib: inflateInit2 failed"};
            plain.resize(std::max<size_t>(1 << 20, map_size * 4));
            size_t in_pos = 0, out_pos = 0;
            zs.avail_in = 0;
            for (;;) {
                if (zs.avail_in == 0) {
                    if (in_pos >= map_size) break;
                    const size_t chunk = std::min<size_t>(map_size - in_pos, 1u << 30);
                    zs.next_in = (Bytef*)(map + in_pos);
                    zs.avail_in = (uInt)chunk;
                    in_pos += chunk;
                }
                if (out_pos == plain.size()) plain.resize(plain.size() * 2);
                const size_t room = std::min<size_t>(plain.size() - out_pos, 1u << 30);
                zs.next_out = (Bytef*)(plain.data() + out_pos);
                zs.avail_out = (uInt)room;
                const int rc = inflate(&zs, Z_NO_FLUSH);
                out_pos += room - zs.avail_out;
                if (rc == Z_STREAM_END) {                  // end of a member: the next one starts right behind it
                    if (zs.avail_in == 0 && in_pos >= map_size) break;
                    if (inflateReset(&zs) != Z_OK) { inflateEnd(&zs); throw Fail{E_IO, "zlib: inflateReset failed"}; }
                    continue;
                }
                if (rc != Z_OK && rc != Z_BUF_ERROR) { inflateEnd(&zs); throw Fail{E_IO, "zlib: corrupt compressed GAMP"}; }
                if (rc == Z_BUF_ERROR && zs.avail_in == 0 && in_pos >= map_size) {
                    inflateEnd(&zs);
                    throw Fail{E_IO, "zlib: truncated compressed GAMP"};
                }
            }
            inflateEnd(&zs);
            plain.resize(out_pos);
            data = plain.data();
            size = plain.size();
        }
        // the records of the stream
        std::vector<std::pair<size_t, size_t>> recs;
        {
            Pb c{(const uint8_t*)data, (const uint8_t*)data + size};
            auto is_tag = [](const uint8_t* b, size_t n) {
                if (n == 0 || n > 24) return false;
                for (size_t i = 0; i < n; ++i)
                    if (!((b[i] >= 'A' && b[i] <= 'Z') || (b[i] >= '0' && b[i] <= '9') || b[i] == '_')) return false;
                return true;
            };
            while (c.more()) {
                const uint64_t count = c.varint();
                for (uint64_t k = 0; k < count; ++k) {
                    Pb m = c.sub();
                    const size_t n = (size_t)(m.e - m.p);
                    if (k == 0 && is_tag(m.p, n)) {      // type tag of the group (a MultipathAlignment never looks like one)
                        if (!(n == 4 && memcmp(m.p, "MGAM", 4) == 0))
                            throw Fail{E_JSON, "not a GAMP stream: type tag '" + std::string((const char*)m.p, n) + "'"};
                        continue;
                    }
                    recs.emplace_back((size_t)((const char*)m.p - data), (size_t)((const char*)m.e - data));
                }
            }
        }
        if (block_records <= 0) block_records = 1 << 20;
        NodeTable nt{node_ids, node_len, n_nodes, {}, {}};
        nt.build();
        const int T = std::max(1, std::min(n_threads, 4096));
        std::vector<Worker> workers((size_t)T);
        std::vector<std::vector<Exc>> exc((size_t)T);
        std::vector<LineOut> lines;
        std::vector<Group> groups;
        size_t pos = 0;
        while (pos < recs.size()) {
            const size_t end = std::min(recs.size(), pos + (size_t)block_records);
            const bool eof = end >= recs.size();
            for (Worker& w : workers) { w.arena.clear(); w.sel.clear(); }
            lines.resize(end - pos);
            run_parallel(T, end - pos, 256, [&](size_t lo, size_t hi, int tid) {
                for (size_t i = lo; i < hi; ++i)
                    decode_line(data + recs[pos + i].first, data + recs[pos + i].second, workers[(size_t)tid], tid, nt, lines[i], true);
            });
            groups.clear();
            BlockTimes bt{0, 0, 0};
            const size_t used = finish_block(D, lines, groups, workers, exc, T, thr, node_len, eof, bt);
            if (used == 0 && !eof) {               // one read group larger than the block: widen the block
                block_records *= 2;
                continue;
            }
            pos += used < lines.size() ? used : end - pos;
        }
    } catch (const Fail& f) {
        D->err_kind = f.kind;
        D->err_msg = f.msg;
    } catch (const std::exception& ex) {
        D->err_kind = E_IO;
        D->err_msg = ex.what();
    }
    if (map) munmap((void*)map, map_size);
    if (fd >= 0) close(fd);
    return D;
}

int sfb_decoded_error(const sfb_decoded* D, const char** msg) {
    if (msg) *msg = D->err_msg.c_str();
    return D->err_kind;
}

int64_t sfb_decoded_size(const sfb_decoded* D, int which) {
    switch (which) {
        case 0: return D->n_groups;
        case 1: return D->n_alns;
        case 2: return D->n_recs;
        case 3: return (int64_t)D->exc_cov.size();
        case 4: return D->n_name_bytes;
        default: return -1;
    }
}

// which: 0 grp_nmates u8[G], 1 grp_aln_off i64[2G+1], 2 mate_seq_len i32[2G], 3 aln_walk_off i64[A+1],
// 4 aln_read_len i32[A], 5 aln_bins u8[5A], 6 aln_stats f64[5A], 7 walk_node i32[R], 8 walk_alen i32[R],
// 9 walk_cov f64[R] (rebuilt from 8 and 10), 10 exc_cov f64[E], 11 name_off i64[2G+1], 12 names bytes
int sfb_decoded_copy(const sfb_decoded* D, int which, void* dst) {
    char* out = (char*)dst;
    auto lead0 = [&] { const int64_t z = 0; memcpy(out, &z, sizeof z); out += sizeof z; };
    auto put = [&](const void* p, size_t bytes) { if (bytes) memcpy(out, p, bytes); out += bytes; };
    if (which == 1 || which == 3 || which == 11) lead0();
    if (which == 10) { put(D->exc_cov.data(), D->exc_cov.size() * sizeof(double)); return 0; }
    if (which < 0 || which > 12) return -1;
    for (const OutBlock& b : D->blocks) {
        switch (which) {
            case 0: put(b.grp_nmates.p, b.grp_nmates.n); break;
            case 1: put(b.grp_aln_off.p, b.grp_aln_off.n * 8); break;
            case 2: put(b.mate_seq_len.p, b.mate_seq_len.n * 4); break;
            case 3: put(b.aln_walk_off.p, b.aln_walk_off.n * 8); break;
            case 4: put(b.aln_read_len.p, b.aln_read_len.n * 4); break;
            case 5: put(b.aln_bins.p, b.aln_bins.n); break;
            case 6: put(b.aln_stats.p, b.aln_stats.n * 8); break;
            case 7: put(b.walk_node.p, b.walk_node.n * 4); break;
            case 8: put(b.walk_alen.p, b.walk_alen.n * 4); break;
            case 9: {
                double* o = (double*)out;
                for (size_t i = 0; i < b.walk_alen.n; ++i) {
                    const int32_t v = b.walk_alen.p[i];
                    o[i] = v < 0 ? D->exc_cov[(size_t)(-(int64_t)v - 1)] : (double)(v & 0xFFFF) / (double)(v >> 16);
                }
                out += b.walk_alen.n * 8;
                break;
            }
            case 11: put(b.name_off.p, b.name_off.n * 8); break;
            case 12: put(b.names.p, b.names.n); break;
        }
    }
    return 0;
}

void sfb_decoded_free(sfb_decoded* D) { delete D; }

}  // extern "C"
