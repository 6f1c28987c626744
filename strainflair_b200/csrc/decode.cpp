// Host-side streaming decoder: `vg view -j -K` multipath-alignment JSON -> the Reads CSR of include/sfb200.h.
//
// Semantics follow the reference line by line where the result depends on it:
//   grouping and best-score filter   get_all_alignments_one_read   json2csv.py:667-739
//   best routes through the subpath DAG   BFS                      json2csv.py:589-665
//   walk / coverage / error ratios   metaalign2align               json2csv.py:551-587
// The reference re-parses ~30 % of the file in its second pass and runs on one thread; here every line is
// parsed once, lines are decoded in parallel, and only the (cheap) grouping is sequential.
#include <algorithm>
#include <chrono>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
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
    while (c.p < c.e && (*c.p == ' ' || *c.p == '\t' || *c.p == '\r' || *c.p == '\n')) ++c.p;
}
[[noreturn]] void bad(const char* what) { throw Fail{E_JSON, std::string("malformed JSON: ") + what}; }

// returns the raw bytes between the quotes (escapes untouched)
inline void str(Cur& c, const char*& b, size_t& n) {
    if (c.p >= c.e || *c.p != '"') bad("expected string");
    b = ++c.p;
    while (c.p < c.e && *c.p != '"') c.p += (*c.p == '\\') ? 2 : 1;
    if (c.p >= c.e) bad("unterminated string");
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
    uint32_t e0, e1;     // range in Line::edits
};
struct Sub {
    uint32_t m0, m1;     // range in Line::maps
    double score;
    bool has_next;
    std::vector<int> next;
    long long to_len;
};
struct Aln {            // one decoded Alignment (json2csv.py:122-134)
    std::vector<int32_t> node;      // dense node index
    std::vector<int32_t> alen;
    std::vector<double> cov;
    int32_t read_len;
    double stats[5];
    uint8_t bins[5];
};
struct Line {
    const char* name; size_t name_n;
    const char* pair; size_t pair_n;
    const char* seq; size_t seq_n;
    bool has_seq, has_subpath;
    std::vector<int> start;
    bool has_start;
    std::vector<Sub> subs;
    std::vector<Mapping> maps;
    std::vector<Edit> edits;
    std::vector<Aln> alns;          // filled by decode_line
    Fail fail{E_NONE, ""};
};

void parse_line(const char* b, const char* e, Line& ln) {
    Cur c{b, e};
    ln.name = ln.pair = ln.seq = nullptr;
    ln.name_n = ln.pair_n = ln.seq_n = 0;
    ln.has_seq = ln.has_subpath = ln.has_start = false;
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
                sp.score = 0.0;
                sp.has_next = false;
                sp.to_len = 0;
                bool has_path = false, has_mapping = false;
                object(c, [&](const char* k2, size_t n2) {
                    if (is(k2, n2, "score")) { ws(c); sp.score = num(c); return true; }
                    if (is(k2, n2, "next")) {
                        sp.has_next = true;
                        array(c, [&] { sp.next.push_back((int)int_like(c)); });
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
                ln.subs.push_back(std::move(sp));
            });
            return true;
        }
        return false;
    });
    if (!has_name) throw Fail{E_KEY, "'name'"};
}

struct NodeTable {
    const int64_t* ids;
    const int32_t* len;
    int64_t n;
    std::vector<int32_t> direct;                 // id -> dense+1 when ids are small
    std::unordered_map<int64_t, int32_t> sparse;
    void build() {
        int64_t mx = -1;
        bool neg = false;
        for (int64_t i = 0; i < n; ++i) { mx = std::max(mx, ids[i]); neg |= ids[i] < 0; }
        if (!neg && mx < 8 * n + 1024) {
            direct.assign((size_t)mx + 1, 0);
            for (int64_t i = 0; i < n; ++i) direct[(size_t)ids[i]] = (int32_t)i + 1;
        } else {
            sparse.reserve((size_t)n * 2);
            for (int64_t i = 0; i < n; ++i) sparse[ids[i]] = (int32_t)i;
        }
    }
    int32_t find(long long id) const {
        if (!direct.empty()) {
            if (id < 0 || (size_t)id >= direct.size() || direct[(size_t)id] == 0) throw Fail{E_KEY, std::to_string(id)};
            return direct[(size_t)id] - 1;
        }
        auto it = sparse.find(id);
        if (it == sparse.end()) throw Fail{E_KEY, std::to_string(id)};
        return it->second;
    }
};

struct Route {
    std::vector<int> path;
    long long read_len;
    double score;
};

inline long long sub_to_len(const Line& ln, int s) {
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
// reference's dict of live routes is insertion ordered; `order` + `slot` reproduce that ordering.
void best_routes(const Line& ln, std::vector<Route>& out) {
    if (!ln.has_start) throw Fail{E_KEY, "'start'"};
    if (!ln.has_seq) throw Fail{E_KEY, "'sequence'"};
    struct Entry {
        int key;
        bool alive;
        std::vector<Route> routes;
    };
    std::vector<Entry> order;
    std::unordered_map<int, int> slot;       // key -> index in order (alive entries only)
    std::deque<int> fifo(ln.start.begin(), ln.start.end());
    for (int s : ln.start) {
        Route r{{s}, sub_to_len(ln, s), ln.subs[s].score};
        auto it = slot.find(s);
        if (it == slot.end()) {
            slot[s] = (int)order.size();
            order.push_back(Entry{s, true, {r}});
        } else {
            order[it->second].routes = {r};
        }
    }
    const long long want = (long long)ln.seq_n;
    while (!fifo.empty()) {
        const int cur = fifo.front();
        fifo.pop_front();
        if (cur < 0 || (size_t)cur >= ln.subs.size()) throw Fail{E_INDEX, "list index out of range"};
        const Sub& sp = ln.subs[cur];
        if (!sp.has_next) continue;
        auto itc = slot.find(cur);
        if (itc == slot.end()) throw Fail{E_KEY, std::to_string(cur)};
        bool all_done = true;
        for (const Route& r : order[itc->second].routes) all_done &= r.read_len == want;
        if (all_done) continue;
        for (int child : sp.next) {
            if (std::find(fifo.begin(), fifo.end(), child) == fifo.end()) fifo.push_back(child);
            const long long c_len = sub_to_len(ln, child);
            const double c_score = ln.subs[child].score;
            auto itk = slot.find(child);
            if (itk == slot.end()) {
                slot[child] = (int)order.size();
                order.push_back(Entry{child, true, {}});
                itk = slot.find(child);
            }
            const int ci = itk->second;
            const int pi = slot.find(cur)->second;
            const size_t n_parent = order[pi].routes.size();      // child == cur would grow it while iterating
            for (size_t q = 0; q < n_parent; ++q) {
                Route nr = order[pi].routes[q];
                nr.path.push_back(child);
                nr.read_len += c_len;
                nr.score += c_score;
                std::vector<Route>& bucket = order[ci].routes;
                if (bucket.empty()) bucket.push_back(std::move(nr));
                else if (nr.score > bucket[0].score) { bucket.clear(); bucket.push_back(std::move(nr)); }
                else if (nr.score == bucket[0].score) bucket.push_back(std::move(nr));
            }
        }
        auto itp = slot.find(cur);
        order[itp->second].alive = false;
        order[itp->second].routes.clear();
        slot.erase(itp);
    }
    size_t alive = 0;
    double top = 0.0;
    bool first = true;
    for (const Entry& en : order)
        if (en.alive) {
            ++alive;
            if (en.routes.empty()) throw Fail{E_INDEX, "list index out of range"};
            if (first || en.routes[0].score > top) top = en.routes[0].score;
            first = false;
        }
    for (const Entry& en : order) {
        if (!en.alive) continue;
        if (alive > 1 && en.routes[0].score < top) continue;
        for (const Route& r : en.routes) out.push_back(r);
    }
}

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

// metaalign2align (json2csv.py:551-587)
void route_to_aln(const Line& ln, const Route& rt, const NodeTable& nt, Aln& a) {
    long long tot_al = 0, tot_match = 0, tot_indel = 0;
    for (int meta : rt.path) {
        const Sub& sp = ln.subs[meta];
        for (uint32_t m = sp.m0; m < sp.m1; ++m) {
            const Mapping& mp = ln.maps[m];
            if (mp.node == -2) throw Fail{E_KEY, "'position'"};
            const int32_t dense = nt.find(mp.node);
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
            if (nt.len[dense] == 0) throw Fail{E_ZERODIV, "division by zero"};
            const double frac = (double)n_al / (double)nt.len[dense];
            if (!a.node.empty() && a.node.back() == dense && a.cov.back() < 1.0) {
                a.cov.back() += frac;
                a.alen.back() += (int32_t)n_al;
            } else {
                a.node.push_back(dense);
                a.cov.push_back(frac);
                a.alen.push_back((int32_t)n_al);
            }
        }
    }
    const double rl = (double)rt.read_len;
    if (rt.read_len == 0 || tot_al == 0) throw Fail{E_ZERODIV, "division by zero"};
    const long long tot_subst = tot_al - tot_match;
    a.read_len = (int32_t)rt.read_len;
    a.stats[0] = rt.score / rl;
    a.stats[1] = (double)tot_match / (double)tot_al;
    a.stats[2] = (double)tot_subst / rl;
    a.stats[3] = (double)tot_indel / rl;
    a.stats[4] = (double)(tot_indel + tot_subst) / rl;
    for (int k = 0; k < 5; ++k) a.bins[k] = bin_of(a.stats[k]);      // here, in the parallel phase (snprintf is slow)
}

void decode_line(Line& ln, const NodeTable& nt) {
    if (!ln.has_subpath) return;
    std::vector<Route> routes;
    best_routes(ln, routes);
    ln.alns.resize(routes.size());
    for (size_t i = 0; i < routes.size(); ++i) route_to_aln(ln, routes[i], nt, ln.alns[i]);
    if (ln.alns.empty()) throw Fail{E_INDEX, "list index out of range"};
}

}  // namespace

struct sfb_decoded {
    std::vector<uint8_t> grp_nmates;
    std::vector<int64_t> grp_aln_off{0};
    std::vector<int32_t> mate_seq_len;
    std::vector<int64_t> aln_walk_off{0};
    std::vector<int32_t> aln_read_len;
    std::vector<uint8_t> aln_bins;
    std::vector<double> aln_stats;
    std::vector<int32_t> walk_node, walk_alen;
    std::vector<double> walk_cov, exc_cov;
    std::vector<int64_t> name_off{0};
    std::string names;
    int err_kind = E_NONE;
    std::string err_msg;
    int64_t n_alns = 0, n_recs = 0;          // running totals (the per-alignment / per-record arrays are filled in parallel)
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

struct MateAcc {
    const char* name; size_t name_n;
    const char* seq; size_t seq_n;
    std::vector<const Aln*> alns;
};

struct EmitItem {       // one retained alignment: where its data goes
    const Aln* a;
    int64_t aln, rec;
};

// serial part of the output: group / mate scalars, names, offsets; the bulky arrays are filled by fill_items
void emit_group(sfb_decoded& D, const std::vector<MateAcc>& mates, std::vector<EmitItem>& items) {
    D.grp_nmates.push_back((uint8_t)mates.size());
    for (int m = 0; m < 2; ++m) {
        if ((size_t)m < mates.size()) {
            const MateAcc& mt = mates[m];
            D.mate_seq_len.push_back((int32_t)mt.seq_n);
            unescape_into(D.names, mt.name, mt.name_n);
            for (const Aln* a : mt.alns) {
                items.push_back(EmitItem{a, D.n_alns, D.n_recs});
                D.n_alns += 1;
                D.n_recs += (int64_t)a->node.size();
                D.aln_walk_off.push_back(D.n_recs);
            }
        } else {
            D.mate_seq_len.push_back(0);
        }
        D.name_off.push_back((int64_t)D.names.size());
        D.grp_aln_off.push_back(D.n_alns);
    }
}

struct Exc {            // a record whose coverage has to be shipped as fp64
    int64_t rec;
    double cov;
};

// parallel part: alignment and record arrays (already sized), packed coverage words, exceptions per thread
void fill_items(sfb_decoded& D, const std::vector<EmitItem>& items, size_t lo, size_t hi, const int32_t* node_len,
                std::vector<Exc>& exc) {
    for (size_t q = lo; q < hi; ++q) {
        const EmitItem& it = items[q];
        const Aln& a = *it.a;
        D.aln_read_len[it.aln] = a.read_len;
        for (int k = 0; k < 5; ++k) {
            D.aln_stats[it.aln * 5 + k] = a.stats[k];
            D.aln_bins[it.aln * 5 + k] = a.bins[k];
        }
        for (size_t i = 0; i < a.node.size(); ++i) {
            const int64_t r = it.rec + (int64_t)i;
            const int32_t al = a.alen[i], nl = node_len[a.node[i]];
            D.walk_node[r] = a.node[i];
            D.walk_cov[r] = a.cov[i];
            // aligned bases | node length << 16 when that reproduces the reference's fp64, else an explicit entry
            if (al >= 0 && al <= 0xFFFF && nl >= 1 && nl <= 0x7FFF && (double)al / (double)nl == a.cov[i]) {
                D.walk_alen[r] = al | (nl << 16);
            } else {
                D.walk_alen[r] = 0;
                exc.push_back(Exc{r, a.cov[i]});
            }
        }
    }
}

// Returns the index of the first line NOT consumed: when `eof` is false the last group of the block may
// continue in the next block, so it is left for the caller to carry over.
size_t group_lines(sfb_decoded& D, std::vector<Line>& lines, double thr, bool eof, std::vector<EmitItem>& items) {
    size_t i = 0;
    const size_t n = lines.size();
    while (i < n) {
        Line& first = lines[i];
        if (first.fail.kind) throw first.fail;
        if (!first.has_seq) throw Fail{E_KEY, "'sequence'"};
        std::vector<MateAcc> mates(1);
        mates[0] = MateAcc{first.name, first.name_n, first.seq, first.seq_n, {}};
        if (first.has_subpath && first.alns[0].stats[0] >= thr)
            for (const Aln& a : first.alns) mates[0].alns.push_back(&a);
        size_t j = i + 1;
        for (; j < n; ++j) {
            Line& ln = lines[j];
            if (ln.fail.kind && ln.fail.kind == E_JSON) throw ln.fail;      // json.loads comes before the name test
            if (!ln.name && !ln.fail.kind) throw Fail{E_KEY, "'name'"};
            if (!same(ln.name, ln.name_n, first.name, first.name_n) &&
                !same(ln.name, ln.name_n, first.pair ? first.pair : "", first.pair_n))
                break;
            if (ln.fail.kind) throw ln.fail;
            if (!ln.has_subpath) continue;
            if (!ln.has_seq) throw Fail{E_KEY, "'sequence'"};
            const bool is_first_seq = same(ln.seq, ln.seq_n, mates[0].seq, mates[0].seq_n);
            if (!is_first_seq && mates.size() == 1) mates.push_back(MateAcc{ln.name, ln.name_n, ln.seq, ln.seq_n, {}});
            MateAcc& tgt = is_first_seq ? mates[0] : mates[1];
            const double best = ln.alns[0].stats[0];
            if (tgt.alns.empty()) {
                if (best >= thr)
                    for (const Aln& a : ln.alns) tgt.alns.push_back(&a);
            } else {
                const double have = tgt.alns[0]->stats[0];
                if (have < best) {
                    tgt.alns.clear();
                    for (const Aln& a : ln.alns) tgt.alns.push_back(&a);
                } else if (have == best) {
                    for (const Aln& a : ln.alns) tgt.alns.push_back(&a);
                }
            }
        }
        if (j == n && !eof) return i;
        emit_group(D, mates, items);
        i = j;
    }
    return n;
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
        double t_par = 0, t_grp = 0, t_fill = 0, t0;
        NodeTable nt{node_ids, node_len, n_nodes, {}, {}};
        nt.build();
        const int T = std::max(1, n_threads);
        std::vector<Line> carry;                 // decoded lines of a group cut by the block boundary (they point into the map)
        std::vector<std::vector<Exc>> exc(T);
        size_t pos = 0;
        bool eof = size == 0;
        while (!eof || !carry.empty()) {
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
            eof = end >= size;
            std::vector<std::pair<size_t, size_t>> spans;
            for (size_t p = pos; p < end;) {
                const void* nl = memchr(data + p, '\n', end - p);
                const size_t q = nl ? (size_t)((const char*)nl - data) : end;
                spans.emplace_back(p, q);
                p = q + 1;
            }
            pos = end;
            t0 = now();
            std::vector<Line> lines(carry.size() + spans.size());
            for (size_t i = 0; i < carry.size(); ++i) lines[i] = std::move(carry[i]);
            const size_t base = carry.size();
            carry.clear();
            auto run_parallel = [&](size_t n_items, size_t grain, auto&& body) {
                const int use = (int)std::max<size_t>(1, std::min<size_t>((size_t)T, n_items / grain + 1));
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
            };
            run_parallel(spans.size(), 256, [&](size_t lo, size_t hi, int) {
                for (size_t i = lo; i < hi; ++i) {
                    try {
                        parse_line(data + spans[i].first, data + spans[i].second, lines[base + i]);
                        decode_line(lines[base + i], nt);
                    } catch (const Fail& f) {
                        lines[base + i].fail = f;
                    }
                }
            });
            t_par += now() - t0;
            t0 = now();
            std::vector<EmitItem> items;
            const size_t used = group_lines(*D, lines, thr, eof, items);
            t_grp += now() - t0;
            t0 = now();
            D->aln_read_len.resize((size_t)D->n_alns);
            D->aln_stats.resize((size_t)D->n_alns * 5);
            D->aln_bins.resize((size_t)D->n_alns * 5);
            D->walk_node.resize((size_t)D->n_recs);
            D->walk_alen.resize((size_t)D->n_recs);
            D->walk_cov.resize((size_t)D->n_recs);
            run_parallel(items.size(), 512, [&](size_t lo, size_t hi, int tid) { fill_items(*D, items, lo, hi, node_len, exc[tid]); });
            for (size_t i = used; i < lines.size(); ++i) carry.push_back(std::move(lines[i]));
            if (eof && used == lines.size()) carry.clear();
            // release the per-line heap blocks from the worker threads (millions of frees, otherwise serial)
            run_parallel(lines.size(), 1024, [&](size_t lo, size_t hi, int) {
                for (size_t i = lo; i < hi; ++i) { Line dead; std::swap(dead, lines[i]); }
            });
            t_fill += now() - t0;
            if (eof && carry.empty()) break;
            if (eof && !carry.empty() && spans.empty()) throw Fail{E_IO, "internal: unconsumed lines at end of file"};
        }
        // explicit coverages in record order (deterministic whatever the thread count)
        std::vector<Exc> all;
        for (auto& v : exc) all.insert(all.end(), v.begin(), v.end());
        std::sort(all.begin(), all.end(), [](const Exc& x, const Exc& y) { return x.rec < y.rec; });
        for (const Exc& e : all) {
            D->exc_cov.push_back(e.cov);
            D->walk_alen[(size_t)e.rec] = -(int32_t)D->exc_cov.size();
        }
        if (trace)
            fprintf(stderr, "[decode] parse+route (parallel) %.3f s, group %.3f s, fill (parallel) %.3f s\n", t_par, t_grp, t_fill);
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

int sfb_decoded_error(const sfb_decoded* D, const char** msg) {
    if (msg) *msg = D->err_msg.c_str();
    return D->err_kind;
}

int64_t sfb_decoded_size(const sfb_decoded* D, int which) {
    switch (which) {
        case 0: return (int64_t)D->grp_nmates.size();
        case 1: return (int64_t)D->aln_read_len.size();
        case 2: return (int64_t)D->walk_node.size();
        case 3: return (int64_t)D->exc_cov.size();
        case 4: return (int64_t)D->names.size();
        default: return -1;
    }
}

// which: 0 grp_nmates u8[G], 1 grp_aln_off i64[2G+1], 2 mate_seq_len i32[2G], 3 aln_walk_off i64[A+1],
// 4 aln_read_len i32[A], 5 aln_bins u8[5A], 6 aln_stats f64[5A], 7 walk_node i32[R], 8 walk_alen i32[R],
// 9 walk_cov f64[R], 10 exc_cov f64[E], 11 name_off i64[2G+1], 12 names bytes
int sfb_decoded_copy(const sfb_decoded* D, int which, void* dst) {
#define SFB_CP(v) memcpy(dst, (v).data(), (v).size() * sizeof((v)[0])); return 0
    switch (which) {
        case 0: SFB_CP(D->grp_nmates);
        case 1: SFB_CP(D->grp_aln_off);
        case 2: SFB_CP(D->mate_seq_len);
        case 3: SFB_CP(D->aln_walk_off);
        case 4: SFB_CP(D->aln_read_len);
        case 5: SFB_CP(D->aln_bins);
        case 6: SFB_CP(D->aln_stats);
        case 7: SFB_CP(D->walk_node);
        case 8: SFB_CP(D->walk_alen);
        case 9: SFB_CP(D->walk_cov);
        case 10: SFB_CP(D->exc_cov);
        case 11: SFB_CP(D->name_off);
        case 12: memcpy(dst, D->names.data(), D->names.size()); return 0;
        default: return -1;
    }
#undef SFB_CP
}

void sfb_decoded_free(sfb_decoded* D) { delete D; }

}  // extern "C"
