// Host-side GFA loader: Pangenome.build_pangenome (reference json2csv.py:223-298) with canonical (:46-61) and
// get_accession_number (:73-81), as arrays.  One pass over a memory-mapped file; S lines give node lengths, P lines
// are put in canonical orientation, merged when their canonical node string is equal (strain copy counts add up) and
// numbered in first-seen order.  The cluster pickle stays with the Python host (it is a Python pickle).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <string>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <unordered_map>
#include <vector>

#include "sfb200.h"

struct sfb_gfa {
    std::vector<int64_t> node_ids;
    std::vector<int32_t> node_len;
    std::vector<int64_t> path_off{0};
    std::vector<int32_t> path_nodes;
    std::vector<int64_t> pstr_off;
    std::vector<int32_t> pstr_id, pstr_cnt, header;
    std::vector<int64_t> strain_off{0};
    std::string strain_names;
    int32_t none_strain = -1;                 // index of the strain that stands for Python's None (Q14), or -1
    std::vector<int64_t> gene_off{0};         // one entry per P line, in file order
    std::string gene_names;
    std::vector<int32_t> gene_pid;
    int err_kind = 0;                         // 0 ok, 2 KeyError, 4 IndexError, 5 I/O, 6 ValueError
    std::string err_msg;
};

namespace {

struct GfaFail {
    int kind;
    std::string msg;
};

inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; }
inline bool is_alpha(char c) { return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z'); }
inline bool is_digit(char c) { return c >= '0' && c <= '9'; }

// int(text) for the plain decimal integers of a GFA
long long to_int(const char* b, const char* e) {
    while (b < e && is_space(*b)) ++b;
    while (e > b && is_space(e[-1])) --e;
    const char* p = b;
    bool neg = false;
    if (p < e && (*p == '-' || *p == '+')) { neg = *p == '-'; ++p; }
    if (p >= e) throw GfaFail{6, "invalid literal for int() with base 10: '" + std::string(b, e) + "'"};
    long long v = 0;
    for (; p < e; ++p) {
        if (!is_digit(*p)) throw GfaFail{6, "invalid literal for int() with base 10: '" + std::string(b, e) + "'"};
        v = v * 10 + (*p - '0');
    }
    return neg ? -v : v;
}

// re.match("[a-zA-Z]+_?[a-zA-Z]*[0-9]+", piece)
bool is_accession(const char* b, const char* e) {
    const char* p = b;
    if (p >= e || !is_alpha(*p)) return false;
    while (p < e && is_alpha(*p)) ++p;
    if (p < e && *p == '_') ++p;
    while (p < e && is_alpha(*p)) ++p;
    return p < e && is_digit(*p);
}

// strain id carried by a path name: drop the last '_' token, first '|' piece that looks like an accession
bool accession_of(const char* b, const char* e, const char*& ab, const char*& ae) {
    const char* last = nullptr;
    for (const char* p = b; p < e; ++p)
        if (*p == '_') last = p;
    const char* stop = last ? last : b;                 // no '_': empty head
    const char* p = b;
    while (p <= stop) {
        const char* q = p;
        while (q < stop && *q != '|') ++q;
        if (is_accession(p, q)) { ab = p; ae = q; return true; }
        if (q >= stop) break;
        p = q + 1;
    }
    return false;
}

void load(sfb_gfa& G, const char* data, size_t size) {
    std::unordered_map<int64_t, int32_t> node_index;
    std::unordered_map<std::string, int32_t> by_content, strains;
    std::vector<std::vector<std::pair<int32_t, int32_t>>> path_strains;   // insertion ordered (strain, copies)
    std::vector<char> in_header;
    std::vector<std::pair<const char*, const char*>> f, toks;
    std::string key;
    node_index.reserve(1 << 20);
    size_t pos = 0;
    while (pos < size) {
        const void* nl = memchr(data + pos, '\n', size - pos);
        const char* b = data + pos;
        const char* e = nl ? (const char*)nl : data + size;
        pos = (size_t)(e - data) + 1;
        while (b < e && is_space(*b)) ++b;              // line.strip()
        while (e > b && is_space(e[-1])) --e;
        f.clear();
        for (const char* p = b;;) {                     // .split('\t')
            const char* q = (const char*)memchr(p, '\t', (size_t)(e - p));
            if (!q) { f.emplace_back(p, e); break; }
            f.emplace_back(p, q);
            p = q + 1;
        }
        const size_t t0 = (size_t)(f[0].second - f[0].first);
        if (t0 != 1) continue;
        const char tag = *f[0].first;
        if (tag == 'S') {
            if (f.size() < 3) throw GfaFail{4, "list index out of range"};
            const int64_t nid = to_int(f[1].first, f[1].second);
            const int32_t len = (int32_t)(f[2].second - f[2].first);
            auto it = node_index.find(nid);
            if (it == node_index.end()) {
                node_index.emplace(nid, (int32_t)G.node_ids.size());
                G.node_ids.push_back(nid);
                G.node_len.push_back(len);
            } else {
                G.node_len[it->second] = len;           // dict overwrite (:251)
            }
        } else if (tag == 'P') {
            if (f.size() < 3) throw GfaFail{4, "list index out of range"};
            toks.clear();
            for (const char* p = f[2].first;;) {        // .split(',')
                const char* q = (const char*)memchr(p, ',', (size_t)(f[2].second - p));
                if (!q) { toks.emplace_back(p, f[2].second); break; }
                toks.emplace_back(p, q);
                p = q + 1;
            }
            // canonical orientation (:46-61); the key keeps the signs, the stored nodes drop them
            bool reversed = false;
            key.clear();
            if (toks.size() == 1) {
                key.assign(toks[0].first, toks[0].second > toks[0].first ? toks[0].second - 1 : toks[0].second);
                key.push_back('+');
            } else {
                const auto& a = toks.front();
                const auto& z = toks.back();
                const long long ia = to_int(a.first, a.second > a.first ? a.second - 1 : a.second);
                const long long iz = to_int(z.first, z.second > z.first ? z.second - 1 : z.second);
                if (ia < iz) {
                    key.assign(f[2].first, f[2].second);
                } else {
                    reversed = true;
                    for (size_t i = toks.size(); i-- > 0;) {
                        const auto& t = toks[i];
                        if (t.second > t.first) {
                            key.append(t.first, t.second - 1);
                            key.push_back(t.second[-1] == '+' ? '-' : '+');
                        } else {
                            throw GfaFail{4, "string index out of range"};
                        }
                        if (i) key.push_back(',');
                    }
                }
            }
            // strain of the gene (first seen over all P lines)
            const char *ab, *ae;
            int32_t sid;
            if (accession_of(f[1].first, f[1].second, ab, ae)) {
                std::string acc(ab, ae);
                auto it = strains.find(acc);
                if (it == strains.end()) {
                    sid = (int32_t)(G.strain_off.size() - 1);
                    strains.emplace(acc, sid);
                    G.strain_names += acc;
                    G.strain_off.push_back((int64_t)G.strain_names.size());
                } else {
                    sid = it->second;
                }
            } else {
                if (G.none_strain < 0) {
                    G.none_strain = (int32_t)(G.strain_off.size() - 1);
                    G.strain_off.push_back((int64_t)G.strain_names.size());
                }
                sid = G.none_strain;
            }
            if ((size_t)sid >= in_header.size()) in_header.resize((size_t)sid + 1, 0);
            G.gene_names.append(f[1].first, f[1].second);
            G.gene_off.push_back((int64_t)G.gene_names.size());
            auto hit = by_content.find(key);
            if (hit != by_content.end()) {              // duplicate content (:260-266)
                auto& d = path_strains[(size_t)hit->second];
                bool found = false;
                for (auto& kv : d)
                    if (kv.first == sid) { ++kv.second; found = true; break; }
                if (!found) d.emplace_back(sid, 1);
                G.gene_pid.push_back(hit->second);
                continue;
            }
            if (!in_header[(size_t)sid]) {              // species_names.add (:268)
                in_header[(size_t)sid] = 1;
                G.header.push_back(sid);
            }
            const int32_t pid = (int32_t)path_strains.size();
            const size_t n = toks.size();
            for (size_t i = 0; i < n; ++i) {
                const auto& t = toks[reversed ? n - 1 - i : i];
                const long long nid = to_int(t.first, t.second > t.first ? t.second - 1 : t.second);
                auto it = node_index.find(nid);
                if (it == node_index.end()) throw GfaFail{2, std::to_string(nid)};     // KeyError (:282)
                G.path_nodes.push_back(it->second);
            }
            G.path_off.push_back((int64_t)G.path_nodes.size());
            path_strains.push_back({{sid, 1}});
            by_content.emplace(key, pid);
            G.gene_pid.push_back(pid);
        }
    }
    G.pstr_off.push_back(0);
    for (const auto& d : path_strains) {
        for (const auto& kv : d) {
            G.pstr_id.push_back(kv.first);
            G.pstr_cnt.push_back(kv.second);
        }
        G.pstr_off.push_back((int64_t)G.pstr_id.size());
    }
}

}  // namespace

extern "C" {

sfb_gfa* sfb_load_gfa(const char* path) {
    sfb_gfa* G = new sfb_gfa();
    int fd = open(path, O_RDONLY);
    if (fd < 0) {
        G->err_kind = 5;
        G->err_msg = std::string("cannot open ") + path;
        return G;
    }
    struct stat sb;
    size_t size = fstat(fd, &sb) == 0 ? (size_t)sb.st_size : 0;
    const char* data = nullptr;
    if (size) {
        void* m = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) {
            G->err_kind = 5;
            G->err_msg = "mmap failed";
            close(fd);
            return G;
        }
        data = (const char*)m;
        madvise(m, size, MADV_SEQUENTIAL);
    }
    try {
        load(*G, data, size);
    } catch (const GfaFail& f) {
        G->err_kind = f.kind;
        G->err_msg = f.msg;
    } catch (const std::exception& ex) {
        G->err_kind = 5;
        G->err_msg = ex.what();
    }
    if (data) munmap((void*)data, size);
    close(fd);
    return G;
}

int sfb_gfa_error(const sfb_gfa* G, const char** msg) {
    if (msg) *msg = G->err_msg.c_str();
    return G->err_kind;
}

// which: 0 nodes, 1 paths, 2 path nodes, 3 (path, strain) entries, 4 strains, 5 bytes of strain names,
// 6 P lines, 7 bytes of gene names, 8 header strains, 9 index of the None strain (-1 if absent)
int64_t sfb_gfa_size(const sfb_gfa* G, int which) {
    switch (which) {
        case 0: return (int64_t)G->node_ids.size();
        case 1: return (int64_t)G->path_off.size() - 1;
        case 2: return (int64_t)G->path_nodes.size();
        case 3: return (int64_t)G->pstr_id.size();
        case 4: return (int64_t)G->strain_off.size() - 1;
        case 5: return (int64_t)G->strain_names.size();
        case 6: return (int64_t)G->gene_pid.size();
        case 7: return (int64_t)G->gene_names.size();
        case 8: return (int64_t)G->header.size();
        case 9: return (int64_t)G->none_strain;
        default: return -1;
    }
}

// which: 0 node_ids i64, 1 node_len i32, 2 path_off i64[P+1], 3 path_nodes i32, 4 pstr_off i64[P+1], 5 pstr_id i32,
// 6 pstr_cnt i32, 7 header i32, 8 strain_off i64[S+1], 9 strain names bytes, 10 gene_off i64[n+1], 11 gene names bytes,
// 12 gene_pid i32[n]
int sfb_gfa_copy(const sfb_gfa* G, int which, void* dst) {
#define SFB_CP(v) memcpy(dst, (v).data(), (v).size() * sizeof((v)[0])); return 0
    switch (which) {
        case 0: SFB_CP(G->node_ids);
        case 1: SFB_CP(G->node_len);
        case 2: SFB_CP(G->path_off);
        case 3: SFB_CP(G->path_nodes);
        case 4: SFB_CP(G->pstr_off);
        case 5: SFB_CP(G->pstr_id);
        case 6: SFB_CP(G->pstr_cnt);
        case 7: SFB_CP(G->header);
        case 8: SFB_CP(G->strain_off);
        case 9: memcpy(dst, G->strain_names.data(), G->strain_names.size()); return 0;
        case 10: SFB_CP(G->gene_off);
        case 11: memcpy(dst, G->gene_names.data(), G->gene_names.size()); return 0;
        case 12: SFB_CP(G->gene_pid);
        default: return -1;
    }
#undef SFB_CP
}

void sfb_gfa_free(sfb_gfa* G) { delete G; }

}  // extern "C"
