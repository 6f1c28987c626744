/* sfb200.h -- C ABI of libsfb200.so, the B200 (sm_100a) abundance engine behind
 * StrainFLAIR's `json2csv` / `compute_strains_abundance` hot path.
 *
 * The reference has no FFI layer: the path sits inside two
 * Python functions.  Each entry point below names the reference code it
 * replaces (file:line relative to the reference checkout).  The Python host
 * (strainflair_b200/engine.py) binds these with ctypes; INTEGRATION.md shows the
 * stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every array is a caller-owned DEVICE pointer (e.g. a torch tensor's
 *     data_ptr()); sizes are int64_t; `stream` is a cudaStream_t passed as void*;
 *   - the library never allocates, frees or synchronises, holds no global
 *     mutable state, and is safe to call concurrently on distinct
 *     streams/buffers;
 *   - every call returns SFB_OK (0) or a negative SFB_E_* code; the message of
 *     the last failure on the calling thread is sfb_last_error().
 *   - accumulating calls (classify, pass1, pass2) ADD into the sfb_accum
 *     arrays: zero them once per job, not per call, so that read shards can be
 *     streamed through the same accumulators.
 */
#ifndef SFB200_H
#define SFB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFB_ABI_VERSION 14

/* ---- status codes ---------------------------------------------------- */
#define SFB_OK            0
#define SFB_E_ARG        -1   /* null pointer / negative size / limit exceeded */
#define SFB_E_CUDA       -2   /* a CUDA runtime call or launch failed           */

/* ---- encodings --------------------------------------------------------- */
/* A match (json2csv.py:341-363 returns (path_id, position)) is a pair of 32-bit words (code, path id):
 *   code bits 0..29  slot = path_off[path_id] + position in the sentinel layout
 *   code bit  30     the reversed walk matched (json2csv.py:862), not the walk itself
 * aln_match holds one pair (x, y) per retained alignment:
 *   x == SFB_MATCH_NONE           no colored path matches
 *   x bit 31 clear                exactly one match: x = code, y = path id
 *   x bit 31 set                  y matches, stored as consecutive (code, path id) pairs of match_buf
 *                                 starting at pair index (x & 0x7fffffff), in the reference's order
 *                                 (forward matches by ascending path id / position, then reverse)
 *   x == SFB_MATCH_SETS           (sfb_match_steps) the matches are two sets of paths of one simple tile, y = first
 *                                 path id p0 of that tile: bit b of aln_sets[2a] = the walk matches path p0 + b, bit b
 *                                 of aln_sets[2a+1] = the reversed walk does; at least one bit is set.  The tuple of a
 *                                 set bit: the occurrence of the walk's first (reversed: last) node on that path, which
 *                                 is occurrence number popcount(node_mask[node] & (bit - 1)) of the node */
#define SFB_SLOT_MASK    0x3fffffffu
#define SFB_MATCH_REV    0x40000000u
#define SFB_MATCH_MULTI  0x80000000u
#define SFB_MATCH_NONE   0xffffffffu
#define SFB_MATCH_SETS   0xfffffffeu
#define SFB_MAX_SLOTS    0x3fffffff          /* n_slots limit (2^30-1)           */

/* Per-mate pass-1 outcome, one nibble per mate in grp_code[g] (mate 0 = low
 * nibble).  Leaves of the decision tree at json2csv.py:830-1053. */
enum {
    SFB_ABSENT = 0,          /* group has no such mate                              */
    SFB_FILTERED = 1,        /* no alignment >= thr            (:834-845, :1013)    */
    SFB_UNASSIGNED = 2,      /* no colored path, not recorded  (:868-910)           */
    SFB_UNASSIGNED_BOOK = 3, /* no colored path, recorded in its cluster (:884,:903,:1037) */
    SFB_DEFER = 4,           /* goes to pass 2                 (:927,:988,:1017,:1028) */
    SFB_ASSIGN_SAME = 5,     /* unique, both mates on one path (:931-959)           */
    SFB_ASSIGN_DIFF = 6,     /* unique after strain reduction  (:991-996)           */
    SFB_MULTI_ALIGN = 7,     /* unresolvable multiple alignments (:935,:948,:998)   */
    SFB_INCONSIST = 8,       /* no common strain               (:1000)              */
    SFB_SE_COUNTED = 9       /* single-end unique: counted, never accumulated (:1047-1050) */
};
/* Per-mate pass-2 outcome in grp_code2[g] (json2csv.py:1085-1274). */
enum {
    SFB_P2_NONE = 0, SFB_P2_ASSIGNED = 1, SFB_P2_MULTI_ALIGN = 2, SFB_P2_SE_ASSIGNED = 3,
    SFB_P2_UNASSIGNED = 4, SFB_P2_INCONSIST = 5
};
/* Index of each module-level counter of json2csv.py:743-790 in sfb_accum.counters. */
enum {
    SFB_C_NB_READS = 0, SFB_C_PAIRS, SFB_C_SINGLES, SFB_C_FILTERED, SFB_C_UNASSIGNED, SFB_C_ASSIGNED_U,
    SFB_C_ASSIGNED_M, SFB_C_INCONSIST, SFB_C_INCONSIST_M, SFB_C_MULTI_ALIGN, SFB_C_SE, SFB_C_SAME_CP,
    SFB_C_DIFF_CP, SFB_C_ONE_NOALIGN, SFB_C_ONE_NOCP, SFB_C_AMBIG_CP_RESOLVED, SFB_C_AMBIG_RESOLVED,
    SFB_C_AMBIG_REDUCED, SFB_C_SE_M, SFB_C_SECOND_PASS,
    SFB_C_BAD_BIN,        /* histogram key outside [0,1] was needed: the reference raises KeyError (Q7) */
    SFB_C_MATCH_OVERFLOW, /* match_buf pairs / app_buf items that did not fit; grow the buffer and rerun */
    SFB_C_NO_PATH,        /* unassigned walk starts on a node no path crosses: IndexError (Q10)    */
    SFB_C_NO_CLUSTER,     /* unassigned read recorded in cluster None: TypeError (Q10)             */
    /* work statistics (feed the roofline arithmetic of bench.py; not part of the reference) */
    SFB_C_ST_CAND,        /* match: candidate occurrences tested                                   */
    SFB_C_ST_COMPARES,    /* match: path-node compares executed                                    */
    SFB_C_ST_TUPLES,      /* match: (path, position) matches found                                 */
    SFB_C_ST_P1_RECORDS,  /* pass 1: alignment-node records added to uniq_abund                    */
    SFB_C_ST_P2_APPLIED,  /* pass 2: (read, candidate path) redistributions                        */
    SFB_C_ST_P2_RECORDS,  /* pass 2: alignment-node records added to mult_abund                    */
    SFB_C_ST_HIST,        /* histogram updates (reads on single-strain paths)                      */
    SFB_N_COUNTERS = 32
};
#define SFB_N_HIST 5          /* mappingscore, identityscore, subst, indel, errors (json2csv.py:335-339) */
#define SFB_N_BINS 101
#define SFB_TABLE_COLS 8      /* columns of sfb_path_reduce's table, see below */

/* ---- graph: Pangenome (json2csv.py:203-319) as arrays ------------------- */
/* Slots: path p occupies slots [path_off[p], path_off[p+1]-1); slot
 * path_off[p+1]-1 is a sentinel (slot_node = -1) so that a walk running off the
 * end of a path fails to compare (json2csv.py:360: the slice is shorter). */
typedef struct sfb_graph {
    int64_t n_nodes, n_paths, n_slots, n_strains, mask_words;
    const int32_t*  slot_node;     /* [n_slots] dense node index, -1 on sentinels            */
    const int32_t*  path_off;      /* [n_paths+1]                                            */
    const int32_t*  occ_off;       /* [n_nodes+1] node -> occurrences (Node.traversed_path + position scan, :352-357) */
    const int32_t*  occ_pair;      /* [2*(n_slots-n_paths)] per occurrence, ascending by slot: (slot, node of
                                      slot+1) -- the successor lets a candidate be rejected without touching
                                      slot_node (8-byte aligned)                                          */
    const int32_t*  blk_path;      /* [2*((n_slots+63)/64)] per block of 64 slots: (path id owning slot 64*b, first
                                      slot inside the block where the next path starts | INT_MAX if none | -1 if
                                      several paths start there): slot -> path with one 8-byte load         */
    const int64_t*  path_seq_len;  /* [n_paths] get_sequence_length (:220-221)               */
    const int32_t*  path_nstrain;  /* [n_paths] len(path.strain_ids)                         */
    const int32_t*  path_strain0;  /* [n_paths] first strain id of the path (histogram row, :334) */
    const uint64_t* path_smask;    /* [n_paths*mask_words] bit s set iff strain s in path.strain_ids */
    const int32_t*  path_cluster;  /* [n_paths] Path.cluster_id, -1 for None                 */
    /* Optional step index (all NULL / 0 without it): get_matching_path as an AND of path sets.  In a "simple" tile
     * (see sfb_tiles: at most 64 paths, no node twice in one path) a walk matches path p at the position of its first
     * node iff p takes every step of the walk, and the reversed walk matches iff p takes every step backwards, so the
     * matches of an alignment are two 64-bit path sets (bit = path id - p0): the AND, over the walk's steps, of the set
     * of paths taking that step.  node_step holds one 32-byte record per dense node:
     *     int32 succ0, succ1     the first two distinct successors of the node on the tile's paths (-1: none)
     *     uint64 mask0, mask1    the paths taking those steps
     *     int32 p0               first path id of the node's tile; -1: the node lies outside the simple tiles (walks
     *                            that start there go through the occurrence-index kernel)
     *     int32 n_succ           distinct successors; those after the first two are step_succ / step_mask[step_off[n]+2 ..)
     * node_mask[n] = the paths through n.  The occurrences of a node are ordered by slot, hence by path: occurrence
     * number k of n belongs to the path of the k-th set bit of node_mask[n]. */
    const void*     node_step;     /* [n_nodes] 32-byte records, 32-byte aligned             */
    const uint64_t* node_mask;     /* [n_nodes]                                              */
    const int32_t*  step_off;      /* [n_nodes+1] distinct successors of node n: step_succ[step_off[n] .. step_off[n+1]) */
    const int32_t*  step_succ;
    const uint64_t* step_mask;
    int64_t         steps_partial; /* != 0: some node with occurrences has p0 = -1           */
} sfb_graph;

/* ---- reads: decoded alignments (json2csv.py:551-739) as CSR -------------- */
typedef struct sfb_reads {
    int64_t n_groups, n_alns, n_records, n_exc;
    const uint8_t* grp_nmates;    /* [G]     1 or 2 (len(mapped_paths), :828)                 */
    const int64_t* grp_aln_off;   /* [2G+1]  retained alignments of (g, mate m): [off[2g+m], off[2g+m+1]) */
    const int32_t* mate_seq_len;  /* [2G]    len(sequence)                                    */
    const int64_t* aln_walk_off;  /* [A+1]                                                    */
    const uint8_t* aln_bins;      /* [A*5]   round(x,2)*100 of the five ratios, 255 = outside [0,1] */
    const int32_t* walk_node;     /* [R]     Alignment.mapped_nodes_id (dense node index)     */
    const int32_t* walk_alen;     /* [R]     >= 0: aligned bases (low 16 bits) | node length << 16, coverage = their
                                             ratio in fp64 (:574); < 0: -(k+1), coverage = exc_cov[k]          */
    const double*  exc_cov;       /* [E]                                                      */
} sfb_reads;

/* ---- compact transfer layout of a read shard ------------------------------
 * What the host copies to the device for the end-to-end path (PCIe is its bound).  Everything travels in the
 * smallest form that the common case allows, with a short list of exceptions beside it:
 *   offsets            as 1-byte counts (prefix sums on the device)
 *   read lengths       one common value + the mates that differ (or 16 bits per mate)
 *   histogram bins     one byte per alignment: index into a dictionary of the shard's frequent 5-tuples
 *   node ids           per alignment the first node as 16 bits above the smallest one of its block of 256 alignments;
 *                      per record the step from the previous node id of the walk in FOUR
 *                      bits (walks follow the graph, whose ids are locally consecutive: a step is +1 or +2 almost always)
 *   aligned bases      one byte for the first and for the last record of a walk; the records in between cover
 *                      their node completely (aligned bases = node length, gathered on the device)
 *   anything else      explicit fp64 coverage of the record (exc_idx / exc_cov): any record may be shipped that way
 * Usable when no mate has more than 255 retained alignments, no walk more than 255 nodes, no read more than
 * 65535 bases.  sfb_expand_reads turns it into the arrays of sfb_reads on the device. */
typedef struct sfb_wire {
    int64_t n_groups, n_alns, n_records, n_exc, n_wide, n_seq_exc, n_bins_exc;
    int64_t seq_len_common;       /* used when mate_seq_len is NULL                   */
    const uint8_t*  grp_nmates;   /* [G]  */
    const uint8_t*  mate_naln;    /* [2G] retained alignments per mate slot           */
    const uint16_t* mate_seq_len; /* [2G] or NULL: seq_len_common for every present mate but the listed ones */
    const int64_t*  seq_exc_idx;  /* [n_seq_exc] mate slot (2 * group + mate)         */
    const int32_t*  seq_exc_val;  /* [n_seq_exc] */
    const uint8_t*  aln_len;      /* [A]  nodes per walk                              */
    const uint8_t*  aln_code;     /* [A]  index into bins_dict, 255 = listed          */
    const uint8_t*  bins_dict;    /* [255*5] */
    const int64_t*  bins_exc_idx; /* [n_bins_exc] alignment                           */
    const uint8_t*  bins_exc_val; /* [n_bins_exc*5] */
    const uint16_t* aln_node0;    /* [A]  first node of the walk minus node0_anchor[a / 256]: read groups in tile order
                                          start within a few hundred node ids of their neighbours (see node0_anchor) */
    const uint8_t*  aln_al_first; /* [A]  aligned bases of the walk's first record, 255 = explicit coverage */
    const uint8_t*  aln_al_last;  /* [A]  ... of its last record (walks of two records or more)            */
    const uint8_t*  walk_step;    /* [(R+1)/2] one nibble per record, record q in bits 4*(q&1).. of byte q/2: node id minus
                                          the previous node id of the walk, + 8 (steps of -7 .. 7; unused for a walk's first
                                          record); 0: the id is wide_node[k] where wide_idx[k] is this record     */
    const int64_t*  exc_idx;      /* [E]  record index of explicit coverage k, ascending */
    const double*   exc_cov;      /* [E]  */
    const int64_t*  wide_idx;     /* [W]  record indices of the wide steps, ascending */
    const int32_t*  wide_node;    /* [W]  */
    const int32_t*  node0_anchor; /* [(A+255)/256] per block of 256 alignments the smallest first node; -1 - k: the block's
                                          first nodes span more than 65535 ids and are node0_wide[256 * k + (a & 255)]   */
    const int32_t*  node0_wide;   /* [256 * number of such blocks] */
} sfb_wire;

/* ---- per-shard work arrays (outputs of match / classify) ----------------- */
typedef struct sfb_work {
    uint32_t* aln_match;          /* [2A]  one (x, y) pair per alignment, see encodings (8-byte aligned) */
    uint64_t* aln_sets;           /* [2A]  path sets of the SFB_MATCH_SETS alignments (sfb_match_steps); may be NULL
                                           when that entry point is not used                   */
    int32_t*  match_buf;          /* [2*match_cap] (code, path id) pairs (8-byte aligned)     */
    int64_t   match_cap;          /* pairs; < 2^31                                            */
    unsigned long long* cursor;   /* [16]  [0] match_buf fill (pairs), [1] deferred groups, [2] app_buf fill,
                                           [3]/[4] groups handed to the generic kernels of classify / pass 2,
                                           [5]/[6] work items taken by sfb_tile_pass1 / sfb_tile_pass2, [7] groups
                                           deferred inside tiles, [8] alignments sfb_match_steps left to the
                                           occurrence-index kernel; zero before the first call of a pass        */
    uint8_t*  grp_code;           /* [G]                                                      */
    uint8_t*  grp_code2;          /* [G]   zero before sfb_pass2                              */
    int32_t*  aln_assign;         /* [A]   >= 0: match code of the unique assignment (:945,:958,:996);
                                           -1: none; <= -2: first alignment of a mate recorded as
                                           unassigned in cluster -(v+2) (:884,:903,:1037)      */
    int32_t*  deferred;           /* [G]   group ids going to pass 2 (unordered)              */
    int32_t*  slow;               /* [G]   scratch: groups outside the register fast path     */
    void*     app_buf;            /* [app_cap] 16-byte pass-2 work items {int32 alignment, uint32 code, double
                                           share}: scratch between the two kernels of sfb_pass2; the number
                                           of matches of a shard (<= A + match_cap) always suffices     */
    int64_t   app_cap;
    int64_t   p2_begin;           /* sfb_classify lists the deferred groups from this one on in `deferred` (and counts the
                                           earlier ones in cursor[7]): the groups before it, staged in tile order, are
                                           resolved by sfb_tile_share, which finds them by their grp_code.  0 without    */
    int64_t   grp_begin, aln_begin; /* first read group / alignment owned by the global-memory kernels (sfb_match,
                                           sfb_classify, sfb_pass1_scatter); the groups before it are in tile order and
                                           belong to sfb_tile_pass1 / sfb_tile_pass2.  0, 0 without tiles             */
} sfb_work;

/* ---- accumulators (Path counters, json2csv.py:113-118, and histograms) ---- */
typedef struct sfb_accum {
    long long* uniq_reads;        /* [P]  total_mapped_unique_reads                           */
    double*    uniq_norm;         /* [P]  total_mapped_unique_reads_normalized                */
    double*    mult_reads;        /* [P]  total_mapped_mult_reads                             */
    double*    mult_norm;         /* [P]  total_mapped_mult_reads_normalized                  */
    unsigned*  mult_hits;         /* [P]  non-zero once pass 2 added to the path (0 <=> the reference prints an int 0) */
    double*    uniq_abund;        /* [n_slots] unique_mapped_abundances                       */
    double*    mult_abund;        /* [n_slots] multiple_mapped_abundances                     */
    unsigned long long* hist;     /* [5*n_strains*101]                                        */
    long long* counters;          /* [SFB_N_COUNTERS]                                         */
    int64_t    det_stride;        /* 0: fp64 RED.ADD straight into the five fp64 arrays (order of the additions, hence
                                     the last bits, vary from run to run).  > 0: order-independent accumulation in
                                     96-bit fixed point -- a term v is the integer floor(v * 2^64); every array has two
                                     64-bit INTEGER planes, plane 0 at x (units of 2^-32), plane 1 at x + det_stride
                                     (units of 2^-64); integer sums are exact, so the result does not depend on thread
                                     order, sharding or GPU count (sum the planes over ranks as int64);
                                     sfb_accum_collapse turns the planes into fp64 before sfb_path_reduce.  Terms and
                                     per-accumulator totals must stay below 2^31                              */
} sfb_accum;

/* ---- tiles: parts of the graph that fit in shared memory ------------------------------------
 * A tile is a set of colored paths with consecutive ids [p0, p1) whose nodes have consecutive dense indices inside
 * [n0, n1), such that no path outside the tile crosses a node of [n0, n1) (a cluster of the pangenome, or a few
 * small neighbouring clusters: StrainFLAIR builds one variation graph per gene cluster and concatenates their id
 * spaces, graphs_construction.py:88, concat_graphs.py).  Every alignment whose first and last node lie in the tile
 * can only match paths of the tile, so a thread block that holds the tile's path nodes, occurrence index and slot
 * accumulators in shared memory runs get_matching_path, the decision tree and the node loops of such read groups
 * without touching the graph or the accumulators in HBM.  Limits per tile: 32767 slots, 255 paths, 65520 nodes,
 * 65535 distinct steps.
 * Read groups are staged in tile order (sfb_tile_permute); the others (mates in different tiles, nodes outside
 * every tile, no alignment at all) follow and go through the global-memory kernels. */
typedef struct sfb_tiles {
    int64_t n_tiles, n_items;
    int64_t max_slots, max_nodes, max_paths, max_edges;  /* of the largest tile: size the shared memory of the kernels */
    int64_t tuple_cap;             /* match tuples per ALIGNMENT kept in shared memory (8, 16 or 32) in tiles that are not
                                      "simple"; aligned pairs with an alignment that has more are handed to the
                                      global-memory kernels through w->slow                                      */
    const int32_t* tile_desc;      /* [8*n_tiles] (p0, p1, n0, n1, flags, 0, 0, 0) per tile; flags bit 0 = "simple":
                                      at most 64 paths and no node twice in one path -- a walk then matches a path
                                      iff the path takes every step of the walk, and get_matching_path is an AND of
                                      the 64-bit path sets of the steps (32-byte aligned)                       */
    const int32_t* items;          /* [4*n_items] (tile, first group, end group, 0): work items of this read shard,
                                      a tile with many groups is cut into several                               */
    /* step index of the simple tiles (zeros elsewhere): the distinct successors of node n on the tile's paths are
       edge_succ[edge_off[n] .. edge_off[n+1]), ascending, edge_mask[e] = the paths that take that step (bit = path id
       - p0); node_mask[n] = the paths through n */
    const int32_t*  edge_off;      /* [n_nodes+1] */
    const int32_t*  edge_succ;     /* [edge_off[n_nodes]] */
    const uint64_t* edge_mask;     /* [edge_off[n_nodes]] */
    const uint64_t* node_mask;     /* [n_nodes] */
} sfb_tiles;

/* ---- entry points ------------------------------------------------------- */
int         sfb_abi_version(void);
const char* sfb_version(void);
const char* sfb_last_error(void);

/* Sizes of the caller-allocated buffers, so that a binding does not restate the formulas:
 *   SFB_WS_EXPANDED      bytes of ONE device allocation that holds, each aligned to 256 bytes and in this order, what
 *                        sfb_expand_reads fills: grp_aln_off[2G+1], aln_walk_off[A+1], its scratch, mate_seq_len[2G],
 *                        aln_bins[5A], walk_node[R], walk_alen[R]
 *   SFB_WS_MATCH_PAIRS   default capacity of match_buf in (code, path id) pairs (the exact need is cursor[0] after a pass)
 *   SFB_WS_APP_ITEMS     capacity of app_buf in 16-byte items that always suffices for that match_buf
 *   SFB_WS_ACCUM_F64     doubles of the fp64 accumulator block: every plane (1, or 2 in the order-independent mode) =
 *                        uniq_norm[P] | mult_reads[P] | mult_norm[P] | uniq_abund[n_slots] | mult_abund[n_slots], plus one
 *                        spare element; det_stride = 3P + 2 n_slots
 *   SFB_WS_ACCUM_I64     64-bit words of the integer block: uniq_reads[P] | hist[5*S*101] | counters[32] | mult_hits (uint32[P])
 * Unused dimensions may be 0.  Returns SFB_E_ARG for an unknown `what` or a negative dimension. */
enum { SFB_WS_EXPANDED = 0, SFB_WS_MATCH_PAIRS = 1, SFB_WS_APP_ITEMS = 2, SFB_WS_ACCUM_F64 = 3, SFB_WS_ACCUM_I64 = 4 };
typedef struct sfb_dims {
    int64_t n_groups, n_alns, n_records, n_paths, n_slots, n_strains, deterministic;
} sfb_dims;
int64_t sfb_workspace_bytes(int what, const sfb_dims* dims);

/* Device-side expansion of the transfer layout.  `dst` is the view the kernels will use: its grp_aln_off[2G+1],
 * mate_seq_len[2G], aln_walk_off[A+1], aln_bins[5A], walk_node[R] and walk_alen[R] point at caller-allocated device
 * buffers that this call fills (prefix sums of the counts, widened lengths, dictionary look-ups, node ids summed
 * along every walk, walk_alen = aligned | node_len[walk_node] << 16 or -(k+1)); its other members are not used.
 * node_len: int32[n_nodes] on the device.  scratch: int64[max(2G, A)/4096 + 2]. */
int sfb_expand_reads(const sfb_wire* w, const int32_t* node_len, const sfb_reads* dst, int64_t* scratch, void* stream);

/* Pangenome.get_matching_path on the walk and, if longer than one node, on the reversed
 * walk (json2csv.py:341-363 with callers :855-866, :1024-1026, :1103-1114, :1231-1233). */
int sfb_match(const sfb_graph* g, const sfb_reads* r, sfb_work* w, sfb_accum* acc, void* stream);
/* The same matches through the step index of `g` (node_step ...), left as path sets: one thread per alignment ANDs the
 * path sets of its walk's steps, both orientations at once, into w->aln_sets and marks the alignment SFB_MATCH_SETS
 * (SFB_MATCH_NONE when both sets are empty).  sfb_classify / sfb_pass2 work on the sets directly (intersections, counts
 * and strain algebra are AND / OR / popcount) and look the slot of a path up only when they assign or share a read.
 * Alignments whose first node lies outside the simple tiles are listed (cursor[8]; the list lives in w->aln_assign,
 * which sfb_classify only writes afterwards) and matched by the occurrence-index kernel of sfb_match, tuples as usual. */
int sfb_match_steps(const sfb_graph* g, const sfb_reads* r, sfb_work* w, sfb_accum* acc, void* stream);

/* Pass-1 decision tree per read group (json2csv.py:828-1053) including the per-read part
 * of fill_abund_dist (:326-327, :333-339): counters, grp_code, aln_assign, deferred list,
 * uniq_reads, uniq_norm, histograms. */
int sfb_classify(const sfb_graph* g, const sfb_reads* r, sfb_work* w, sfb_accum* acc, void* stream);

/* Node-level part of fill_abund_dist (json2csv.py:329-330) for every alignment with
 * aln_assign >= 0: uniq_abund[slot+i] += cov[i]. */
int sfb_pass1_scatter(const sfb_graph* g, const sfb_reads* r, const sfb_work* w, sfb_accum* acc, void* stream);

/* Pass 2 (json2csv.py:1085-1274) over the deferred groups.  `uniq_reads_global` is the
 * per-path unique count summed over ALL read shards (acc->uniq_reads when there is one). */
int sfb_pass2(const sfb_graph* g, const sfb_reads* r, sfb_work* w, sfb_accum* acc,
              const long long* uniq_reads_global, void* stream);

/* The two halves of sfb_pass2, for callers that want to time or overlap them: candidate selection, ratios and
 * per-path counters (json2csv.py:1127-1156 and twins) into w->app_buf, then the node loops (:1157-1158). */
int sfb_pass2_resolve(const sfb_graph* g, const sfb_reads* r, sfb_work* w, sfb_accum* acc,
                      const long long* uniq_reads_global, void* stream);
int sfb_pass2_scatter(const sfb_graph* g, const sfb_reads* r, sfb_work* w, sfb_accum* acc, void* stream);

/* Tile kernels: match + pass-1 decision tree + per-read and node-level parts of fill_abund_dist in one kernel
 * (json2csv.py:341-363, :828-1053, :321-339) for the read groups of t->items; then, once uniq_reads is complete over
 * all read shards, pass 2 for their deferred groups (json2csv.py:1085-1274; matches are recomputed from shared memory
 * as the reference recomputes them from the file).  Outputs as the global-memory kernels': grp_code, grp_code2,
 * aln_assign, the accumulators (shared-memory bins in 96-bit fixed point, flushed once per work item), counters,
 * histograms.  Aligned pairs with more than t->tuple_cap match tuples in a mate get their match lists written to
 * aln_match / match_buf and are appended to w->slow: call sfb_classify afterwards (it classifies them, adds their
 * node loops itself and leaves their deferred ones to sfb_pass2). */
int sfb_tile_pass1(const sfb_graph* g, const sfb_tiles* t, const sfb_reads* r, sfb_work* w, sfb_accum* acc, void* stream);
int sfb_tile_pass2(const sfb_graph* g, const sfb_tiles* t, const sfb_reads* r, sfb_work* w, sfb_accum* acc,
                   const long long* uniq_reads_global, void* stream);
/* Node-level sums in shared memory on top of the path-set kernels (sfb_match_steps, sfb_classify), for the read groups of
 * t->items (staged in tile order): the slot accumulators of the item's tile are 96-bit fixed-point bins in shared memory,
 * flushed once per item -- 32-bit shared-memory atomics instead of 64-bit REDs at the L2, whose rate bounds
 * sfb_pass1_scatter / sfb_pass2_scatter.  Only tile_desc, items, max_slots and max_paths of `t` are used.
 *   sfb_tile_scatter1  the node loop of fill_abund_dist (json2csv.py:329-330) for the alignments of the items' groups with
 *                      aln_assign >= 0 (call sfb_pass1_scatter with w->aln_begin = first alignment after the items for
 *                      the rest); takes its items through cursor[5]
 *   sfb_tile_share     pass 2 (json2csv.py:1085-1274) for the items' groups that sfb_classify deferred (w->p2_begin = first
 *                      group after the items): candidate selection and ratios on the path sets, per-path sums and node
 *                      loops into the bins; a deferred group that holds tuple lists is appended to w->deferred for
 *                      sfb_pass2_resolve, which must run afterwards; takes its items through cursor[6] */
int sfb_tile_scatter1(const sfb_graph* g, const sfb_tiles* t, const sfb_reads* r, sfb_work* w, sfb_accum* acc, void* stream);
int sfb_tile_share(const sfb_graph* g, const sfb_tiles* t, const sfb_reads* r, sfb_work* w, sfb_accum* acc,
                   const long long* uniq_reads_global, void* stream);
/* Dynamic shared memory (bytes) the tile kernels need for these tiles; SFB_E_ARG when a tile exceeds the limits. */
int64_t sfb_tile_smem_bytes(const sfb_tiles* t);

/* Order-independent mode only (acc->det_stride > 0): x[i] = plane0[i] * 2^-32 + plane1[i] * 2^-64 (fp64) for the n
 * accumulators at x. */
int sfb_accum_collapse(double* x, int64_t n, int64_t det_stride, void* stream);

/* A slot array (uniq_abund / mult_abund: one sentinel slot after every path) without its sentinels, path after path:
 * plain[n_slots - n_paths] = the per-node sums Path.unique_mapped_abundances / multiple_mapped_abundances hold in the
 * reference (json2csv.py:107-108), for the read-back of the whole arrays. */
int sfb_drop_sentinels(const sfb_graph* g, const double* slots, double* plain, void* stream);

/* Pangenome.print_gene_table's per-path reductions (json2csv.py:440-469) for paths
 * [p_begin, p_end): table[(p-p_begin)*8 + k], k = 0 nb_uniq_mapped_normalized,
 * 1 nb_multimapped, 2 nb_multimapped_normalized, 3 mean_abund_uniq, 4 mean_abund_uniq_nz,
 * 5 mean_abund_multiple, 6 mean_abund_multiple_nz, 7 ratio_covered_nodes. */
int sfb_path_reduce(const sfb_graph* g, const sfb_accum* acc, int64_t p_begin, int64_t p_end,
                    double* table, void* stream);

/* print_unassigned_info's per-cluster strain copy counts (json2csv.py:544):
 * out[c*n_strains + s] = sum over clu_paths[clu_off[c]..clu_off[c+1]) of the copy count of s. */
int sfb_cluster_strain_counts(int64_t n_clusters, int64_t n_strains, const int32_t* clu_off,
                              const int32_t* clu_paths, const int32_t* pstr_off, const int32_t* pstr_id,
                              const int32_t* pstr_cnt, long long* out, void* stream);

/* compute_strains_abundance (compute_strains_abundance.py:52-69).
 * Inputs per gene row p: CSR of its positive strain counts (row_off/row_strain/row_count) and
 * mean_abund_multiple, mean_abund_multiple_nz, ratio_covered_nodes (NaN already 0, :50).
 * ws_rows int32[2P], ws_vals double[2P], ws_counts double[4S] are scratch.
 * out[s*5 + k]: 0 detected_genes (NaN when 0/0), 1 mean_abund, 2 mean_abund_nz,
 * 3 median_abund, 4 median_abund_nz, thresholded and normalised to 100. */
int sfb_strain_aggregate(int64_t n_rows, int64_t n_strains, const int32_t* row_off,
                         const int32_t* row_strain, const double* row_count, const double* mam,
                         const double* mam_nz, const double* ratio_cov, double thr, int32_t* ws_rows,
                         double* ws_vals, double* ws_counts, double* out, void* stream);

/* ---- host-side alignment decoder (no GPU involved) ------------------------------------------
 * get_all_alignments_one_read + BFS + metaalign2align (json2csv.py:551-739) over a whole
 * `vg view -j -K` JSON file, multi-threaded, one parse per line (the reference parses deferred reads
 * twice).  node_ids[k] is the GFA id of dense node k, node_len[k] its length.  The handle owns the
 * Reads arrays of sfb_reads plus names/statistics; sizes: which = 0 groups, 1 alignments, 2 records,
 * 3 explicit coverages, 4 bytes of names.  sfb_decoded_copy(which): 0 grp_nmates u8[G],
 * 1 grp_aln_off i64[2G+1], 2 mate_seq_len i32[2G], 3 aln_walk_off i64[A+1], 4 aln_read_len i32[A],
 * 5 aln_bins u8[5A], 6 aln_stats f64[5A], 7 walk_node i32[R], 8 walk_alen i32[R], 9 walk_cov f64[R],
 * 10 exc_cov f64[E], 11 name_off i64[2G+1], 12 names bytes.
 * sfb_decoded_error: 0 ok, 1 malformed JSON (ValueError), 2 KeyError, 3 ZeroDivisionError,
 * 4 IndexError, 5 I/O, 6 TypeError/ValueError of int() -- the exception the reference raises. */
typedef struct sfb_decoded sfb_decoded;
sfb_decoded* sfb_decode_json(const char* path, const int64_t* node_ids, const int32_t* node_len, int64_t n_nodes,
                             double thr, int n_threads, int64_t block_bytes);
/* The same from the GAMP file `vg mpmap` writes (StrainFLAIR.sh:582), without the `vg view -j -K` conversion to JSON
 * (:592-602): vg's stream framing (groups of length-prefixed messages, optional "MGAM" type tag, optional BGZF / gzip
 * compression) around MultipathAlignment protobuf messages.  Restated from the published format -- no vg binary or GAMP
 * file is at hand to pin it; block_records: messages decoded per block (<= 0: default). */
sfb_decoded* sfb_decode_gamp(const char* path, const int64_t* node_ids, const int32_t* node_len, int64_t n_nodes,
                             double thr, int n_threads, int64_t block_records);
int          sfb_decoded_error(const sfb_decoded* d, const char** msg);
int64_t      sfb_decoded_size(const sfb_decoded* d, int which);
int          sfb_decoded_copy(const sfb_decoded* d, int which, void* dst);
void         sfb_decoded_free(sfb_decoded* d);

/* ---- host-side packer of the transfer layout (no GPU involved) ----------------------------
 * sfb_wire_pack: `host_reads` holds HOST pointers; returns the arrays of sfb_wire for that shard (multi-threaded, the
 * result does not depend on n_threads).  sfb_packed_status: 0 ok, 1 the layout does not apply (a mate with more than
 * 255 alignments, a walk of more than 255 nodes, a read of more than 65535 bases, or so many exceptions that the plain
 * arrays are smaller): ship the sfb_reads arrays as they are; 5 out of memory.  sfb_packed_size(which): bytes of array
 * `which` (0..17, the order of the pointers in sfb_wire; 0 bytes = NULL pointer), 100: seq_len_common. */
typedef struct sfb_packed sfb_packed;
sfb_packed*  sfb_wire_pack(const sfb_reads* host_reads, int n_threads);
int          sfb_packed_status(const sfb_packed* p, const char** msg);
int64_t      sfb_packed_size(const sfb_packed* p, int which);
int          sfb_packed_copy(const sfb_packed* p, int which, void* dst);
void         sfb_packed_free(sfb_packed* p);

/* ---- host-side locality order of the read groups (no GPU involved) ------------------------
 * The sums of the abundance pass do not depend on the order of the read groups; staged stable by the first node of their
 * first retained alignment (groups without one last), neighbouring groups gather the same index entries, paths and
 * accumulator slots.  `src` and `dst` hold HOST pointers; dst's grp_nmates[G], grp_aln_off[2G+1], mate_seq_len[2G],
 * aln_walk_off[A+1], aln_bins[5A], walk_node[R], walk_alen[R] are caller-allocated and filled with the permuted CSR
 * (exc_cov is shared: explicit-coverage records keep their index).  order[G]: old index of new group k; aln_index[A]: old
 * index of new alignment k (per-alignment side arrays and outputs follow it).  The result does not depend on n_threads. */
int sfb_locality_permute(const sfb_reads* src, int n_threads, int64_t* order, int64_t* aln_index, const sfb_reads* dst);

/* Tile order (see sfb_tiles): as sfb_locality_permute, but the key of a group is the tile that holds the first and
 * the last node of every one of its alignments (node_tile[n] = tile of dense node n, -1 = none); groups without such
 * a tile -- and groups without any alignment, or with more than 255 alignments in a mate -- go last, in file order.  tile_grp_off[n_tiles + 2]: the groups of tile
 * t are [off[t], off[t+1]) of the permuted order -- inside a tile stable by the first node of the group's first alignment,
 * so that neighbouring groups touch neighbouring positions of the tile's paths (the tiles must be numbered in node order:
 * every node of tile t before every node of tile t+1) -- the others [off[n_tiles], off[n_tiles+1] = G). */
int sfb_tile_permute(const sfb_reads* src, const int32_t* node_tile, int64_t n_tiles, int n_threads, int64_t* order,
                     int64_t* aln_index, const sfb_reads* dst, int64_t* tile_grp_off);

/* ---- host-side analysis of the unassigned reads (no GPU involved) ---------------------------
 * Cluster.min_strains_inference (json2csv.py:149-201, with canonical_tiny :63-71, check_incomp_comp :83-97, sublist
 * :99-100) for every cluster: the walks of cluster c are walks clu_off[c] .. clu_off[c+1]-1, walk w is
 * nodes[walk_off[w] .. walk_off[w+1]-1] (node ids as in the GFA).  Per cluster: size of the largest set of mutually
 * incompatible walks, number of walks kept by the support filter, the same size among the kept walks, mean number of
 * kept walks per node (NaN when none is kept); zeros for a cluster without walks.  Multi-threaded over clusters. */
int sfb_min_strains(int64_t n_clusters, const int64_t* clu_off, const int64_t* walk_off, const int64_t* nodes,
                    int n_threads, int64_t* before, int64_t* kept, int64_t* after, double* support);

/* ---- host-side GFA loader (no GPU involved) ------------------------------------------------
 * Pangenome.build_pangenome + canonical + get_accession_number (json2csv.py:46-81, 223-298) over a memory-mapped
 * GFA.  sfb_gfa_size(which): 0 nodes, 1 paths, 2 path nodes, 3 (path, strain) entries, 4 strains, 5 bytes of strain
 * names, 6 P lines, 7 bytes of gene names, 8 header strains, 9 index of the strain standing for None (-1 if none).
 * sfb_gfa_copy(which): 0 node_ids i64, 1 node_len i32, 2 path_off i64[P+1], 3 path_nodes i32, 4 pstr_off i64[P+1],
 * 5 pstr_id i32, 6 pstr_cnt i32, 7 header i32, 8 strain_off i64[S+1], 9 strain names, 10 gene_off i64[n+1],
 * 11 gene names, 12 gene_pid i32[n] (P line -> path id).  sfb_gfa_error: 0 ok, 2 KeyError (path node without S line),
 * 4 IndexError (short line), 5 I/O, 6 ValueError (int()). */
typedef struct sfb_gfa sfb_gfa;
sfb_gfa*     sfb_load_gfa(const char* path);
int          sfb_gfa_error(const sfb_gfa* g, const char** msg);
int64_t      sfb_gfa_size(const sfb_gfa* g, int which);
int          sfb_gfa_copy(const sfb_gfa* g, int which, void* dst);
void         sfb_gfa_free(sfb_gfa* g);

#ifdef __cplusplus
}
#endif
#endif /* SFB200_H */
