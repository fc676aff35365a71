// See mutation_annotated_tree.hpp.  Behaviour follows reference src/mutation_annotated_tree.cpp (line ranges
// cited per function); the code is an independent implementation.
#include "mutation_annotated_tree.hpp"
#include "flat_mat.hpp"
#include "usher_b200.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <chrono>
#include <thread>
#include <array>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <functional>
#include <sstream>

#include "usher_graph.hpp"

namespace Mutation_Annotated_Tree {

// ---------------------------------------------------------------- nucleotide codes (reference :17-208)
int8_t get_nuc_id(char c) {
    switch (c) {
        case 'a': case 'A': return 1;
        case 'c': case 'C': return 2;
        case 'g': case 'G': return 4;
        case 't': case 'T': return 8;
        case 'R': return 5;
        case 'Y': return 10;
        case 'S': return 6;
        case 'W': return 9;
        case 'K': return 12;
        case 'M': return 3;
        case 'B': return 14;
        case 'D': return 13;
        case 'H': return 11;
        // 'V' falls through to N in the reference (missing break, :65-71); kept for parity
        default: return 15;
    }
}
int8_t get_nuc_id(const std::vector<int8_t>& v) {
    int8_t r = 0;
    for (auto n : v) r = (int8_t)(r + (1 << n));
    return r;
}
char get_nuc(int8_t id) {
    static const char tab[16] = {'N', 'A', 'C', 'M', 'G', 'R', 'S', 'V', 'T', 'W', 'Y', 'H', 'K', 'D', 'B', 'N'};
    return (id >= 1 && id <= 14) ? tab[id] : 'N';
}
int8_t get_nt(int8_t id) {
    switch (id) { case 1: return 0; case 2: return 1; case 4: return 2; case 8: return 3; default: return -1; }
}
std::vector<int8_t> get_nuc_vec_from_id(int8_t id) {
    // goes through the IUPAC letter like the reference (get_nuc then get_nuc_vec): 7 -> 'V' -> {0,1,2}
    const int8_t bits = (id >= 1 && id <= 14) ? id : 15;
    std::vector<int8_t> v;
    for (int8_t b = 0; b < 4; b++) if (bits & (1 << b)) v.push_back(b);
    return v;
}

// ---------------------------------------------------------------- Node
void Node::add_mutation(const Mutation& mut) {
    auto it = std::lower_bound(mutations.begin(), mutations.end(), mut);
    if (it != mutations.end() && it->position == mut.position) {
        if (it->par_nuc != mut.mut_nuc) {
            it->mut_nuc = mut.mut_nuc;
        } else {
            if (it->mut_nuc != mut.par_nuc) {
                fprintf(stderr, "ERROR: add_mutation: consecutive mutations at same position disagree on nuc (%s > %s) "
                        "-- called out of order?\n", it->get_string().c_str(), mut.get_string().c_str());
                exit(1);
            }
            const int pos = it->position;
            mutations.erase(std::remove_if(mutations.begin(), mutations.end(),
                                           [pos](const Mutation& m) { return m.position == pos; }),
                            mutations.end());
        }
    } else {
        mutations.insert(it, mut);
    }
}

// ---------------------------------------------------------------- Tree
Tree& Tree::operator=(Tree&& o) noexcept {
    if (this != &o) {
        for (auto& kv : all_nodes) delete kv.second;
        root = o.root; o.root = nullptr;
        condensed_nodes = std::move(o.condensed_nodes);
        condensed_order = std::move(o.condensed_order);
        condensed_leaves = std::move(o.condensed_leaves);
        curr_internal_node = o.curr_internal_node;
        all_nodes = std::move(o.all_nodes);
        o.all_nodes.clear();
    }
    return *this;
}
Tree::~Tree() { for (auto& kv : all_nodes) delete kv.second; }

Node* Tree::create_node(const std::string& id, float len, size_t num_annotations) {
    for (auto& kv : all_nodes) delete kv.second;
    all_nodes.clear();
    Node* n = new Node();
    n->identifier = id;
    n->level = 1;
    n->branch_length = len;
    n->clade_annotations.assign(num_annotations, "");
    root = n;
    all_nodes[id] = n;
    return n;
}
Node* Tree::create_node(const std::string& id, Node* par, float len) {
    auto ins = all_nodes.emplace(id, nullptr);   // one hash lookup for the duplicate check and the insertion
    if (!ins.second) {
        fprintf(stderr, "Error: %s already in the tree!\n", id.c_str());
        exit(1);
    }
    Node* n = new Node();
    n->identifier = id;
    n->parent = par;
    n->level = par->level + 1;
    n->branch_length = len;
    n->clade_annotations.assign(get_num_annotations(), "");
    ins.first->second = n;
    par->children.push_back(n);
    return n;
}
Node* Tree::create_node(const std::string& id, const std::string& parent_id, float len) {
    return create_node(id, all_nodes.at(parent_id), len);
}
Node* Tree::get_node(const std::string& id) const {
    auto it = all_nodes.find(id);
    return it == all_nodes.end() ? nullptr : it->second;
}
std::vector<Node*> Tree::rsearch(const std::string& nid, bool include_self) const {
    std::vector<Node*> anc;
    Node* n = get_node(nid);
    if (!n) return anc;
    if (include_self) anc.push_back(n);
    for (n = n->parent; n; n = n->parent) anc.push_back(n);
    return anc;
}
std::string Tree::get_clade_assignment(const Node* n, int clade_id, bool include_self) const {
    for (auto a : rsearch(n->identifier, include_self))
        if ((int)a->clade_annotations.size() > clade_id && a->clade_annotations[clade_id] != "")
            return a->clade_annotations[clade_id];
    return "UNDEFINED";
}
size_t Tree::get_num_leaves(Node* node) const {   // childless nodes below `node`; a condensed node counts 1 (:866-879)
    if (!node) node = root;
    size_t cnt = 0;
    std::vector<Node*> st{node};
    while (!st.empty()) {
        Node* u = st.back();
        st.pop_back();
        if (u->children.empty()) cnt++;
        for (auto c : u->children) st.push_back(c);
    }
    return cnt;
}
std::vector<Node*> Tree::get_leaves(const std::string& nid) const {
    std::vector<Node*> out;
    Node* start = nid.empty() ? root : get_node(nid);
    if (!start) return out;
    std::deque<Node*> q{start};
    while (!q.empty()) {
        Node* u = q.front();
        q.pop_front();
        if (u->children.empty()) out.push_back(u);
        for (auto c : u->children) q.push_back(c);
    }
    return out;
}
std::vector<Node*> Tree::breadth_first_expansion(const std::string& nid) const {   // :1225-1251
    std::vector<Node*> out;
    Node* start = nid.empty() ? root : get_node(nid);
    if (!start) return out;
    out.push_back(start);
    for (size_t h = 0; h < out.size(); h++)
        for (auto c : out[h]->children) out.push_back(c);
    return out;
}
std::vector<Node*> Tree::depth_first_expansion(Node* node) const {                 // :1253-1274, pre-order
    std::vector<Node*> out;
    if (!node) node = root;
    if (!node) return out;
    std::vector<Node*> st{node};
    while (!st.empty()) {
        Node* u = st.back();
        st.pop_back();
        out.push_back(u);
        for (auto it = u->children.rbegin(); it != u->children.rend(); ++it) st.push_back(*it);
    }
    return out;
}
size_t Tree::get_parsimony_score() const {
    size_t s = 0;
    for (auto n : depth_first_expansion()) s += n->mutations.size();
    return s;
}

static void relevel(Node* n) {
    std::vector<Node*> st{n};
    while (!st.empty()) {
        Node* u = st.back();
        st.pop_back();
        u->level = u->parent ? u->parent->level + 1 : 1;
        for (auto c : u->children) st.push_back(c);
    }
}

// Reference :1135-1222.  The placement path only ever moves the chosen node under a brand-new internal node
// whose single child (the new sample) has no mutations yet, i.e. the "simple re-link" case; the merge cases
// of the reference (destination already has a child with the same non-empty mutation set) are reported.
void Tree::move_node(const std::string& source_id, const std::string& dest_id) {
    Node* src = all_nodes.at(source_id);
    Node* dst = all_nodes.at(dest_id);
    Node* old = src->parent;
    if (old == dst) {
        fprintf(stderr, "ERROR: move_node: dest_id=%s but that is already parent of source_id=%s\n", dest_id.c_str(),
                source_id.c_str());
        exit(1);
    }
    if (!src->mutations.empty()) {
        for (auto c : dst->children) {
            if (c == old || c->mutations.size() != src->mutations.size()) continue;
            bool same = true;
            for (size_t i = 0; i < c->mutations.size() && same; i++)
                same = c->mutations[i].position == src->mutations[i].position &&
                       c->mutations[i].mut_nuc == src->mutations[i].mut_nuc;
            if (same) {
                fprintf(stderr, "ERROR: move_node: merging with an identical sibling is not supported by this build\n");
                exit(1);
            }
        }
    }
    src->parent = dst;
    src->branch_length = -1.0f;
    dst->children.push_back(src);
    old->children.erase(std::find(old->children.begin(), old->children.end(), src));
    if (old->children.empty()) remove_node(old->identifier, true);
    relevel(src);
}

// Remove a childless node (and, like the reference :1002-1061, any ancestor left childless by it).
void Tree::remove_node(const std::string& nid, bool move_level) {
    (void)move_level;
    Node* n = get_node(nid);
    while (n) {
        Node* par = n->parent;
        if (par) par->children.erase(std::find(par->children.begin(), par->children.end(), n));
        else root = nullptr;
        all_nodes.erase(n->identifier);
        delete n;
        n = (par && par->children.empty()) ? par : nullptr;
    }
}

// Sibling leaves without mutations collapse into node_<k>_condensed_<n>_leaves appended to the parent
// (reference :1287-1332).
void Tree::condense_leaves(const std::vector<std::string>& missing) {
    if (!condensed_nodes.empty()) {
        fprintf(stderr, "WARNING: tree contains condensed nodes. Uncondensing fist.\n");
        uncondense_leaves();
    }
    auto is_missing = [&](const std::string& s) { return std::find(missing.begin(), missing.end(), s) != missing.end(); };
    std::vector<std::string> leaf_ids;
    for (auto l : get_leaves()) leaf_ids.push_back(l->identifier);
    for (auto& id : leaf_ids) {
        Node* l1 = get_node(id);
        if (!l1 || is_missing(id) || !l1->mutations.empty() || !l1->parent) continue;
        std::vector<Node*> poly;
        for (auto l2 : l1->parent->children)
            if (!is_missing(l2->identifier) && l2->is_leaf() && l2->mutations.empty()) poly.push_back(l2);
        if (poly.size() > 1) {
            const std::string name = "node_" + std::to_string(1 + condensed_nodes.size()) + "_condensed_" +
                                     std::to_string(poly.size()) + "_leaves";
            create_node(name, l1->parent, l1->branch_length);
            std::vector<std::string> members;
            for (auto p : poly) members.push_back(p->identifier);
            for (auto& m : members) {
                condensed_leaves.insert(m);
                Node* p = get_node(m);
                p->parent->children.erase(std::find(p->parent->children.begin(), p->parent->children.end(), p));
                all_nodes.erase(m);
                delete p;
            }
            condensed_nodes[name] = members;
            condensed_order.push_back(name);
        }
    }
}

void Tree::uncondense_leaves() {   // reference :1334-1383
    std::vector<std::string> order = condensed_order;
    for (auto& kv : condensed_nodes)
        if (std::find(order.begin(), order.end(), kv.first) == order.end()) order.push_back(kv.first);
    for (auto& name : order) {
        auto it = condensed_nodes.find(name);
        if (it == condensed_nodes.end()) continue;
        Node* n = get_node(name);
        if (!n) continue;
        Node* par = n->parent ? n->parent : n;
        const auto& members = it->second;
        const size_t k = members.size();
        auto add = [&](const std::string& id, Node* p, float len) {
            Node* nn = new Node();
            nn->identifier = id; nn->parent = p; nn->level = p->level + 1; nn->branch_length = len;
            nn->clade_annotations.assign(get_num_annotations(), "");
            all_nodes[id] = nn;
            p->children.push_back(nn);
        };
        if (k > 1 && !n->mutations.empty()) {
            all_nodes.erase(n->identifier);
            n->identifier = new_internal_node_id();
            all_nodes[n->identifier] = n;
            for (auto& m : members) add(m, n, -1.0f);
        } else if (k > 1) {
            all_nodes.erase(n->identifier);
            n->identifier = members[0];
            all_nodes[n->identifier] = n;
            for (size_t s = 1; s < k; s++) add(members[s], par, n->branch_length);
        } else if (k == 1) {
            all_nodes.erase(n->identifier);
            n->identifier = members[0];
            all_nodes[n->identifier] = n;
        }
    }
    condensed_nodes.clear();
    condensed_order.clear();
    condensed_leaves.clear();
}

// ---------------------------------------------------------------- newick
// Writer: plain newick of the subtree, branch length = number of mutations on the branch (the reference forces
// this, :228-230), internal names only when asked, a condensed leaf optionally expanded to its members.
static void write_subtree(std::ostringstream& ss, const Tree& T, Node* top, bool names, bool lens, bool expand) {
    struct Frame { Node* n; size_t next; };
    std::vector<Frame> st{{top, 0}};
    auto emit_len = [&](Node* n) { if (lens) { ss << ':' << static_cast<float>(n->mutations.size()); } };
    while (!st.empty()) {
        Frame& f = st.back();
        Node* n = f.n;
        if (n->children.empty()) {
            auto cn = expand ? T.condensed_nodes.find(n->identifier) : T.condensed_nodes.end();
            if (cn != T.condensed_nodes.end()) {
                for (size_t i = 0; i < cn->second.size(); i++) { if (i) ss << ','; ss << cn->second[i]; }
            } else {
                ss << n->identifier;
            }
            emit_len(n);
            st.pop_back();
            continue;
        }
        if (f.next == 0) ss << '(';
        if (f.next < n->children.size()) {
            if (f.next) ss << ',';
            Node* c = n->children[f.next++];
            st.push_back({c, 0});
            continue;
        }
        ss << ')';
        if (names) ss << n->identifier;
        emit_len(n);
        st.pop_back();
    }
    ss << ';';
}
std::string get_newick_string(const Tree& T, Node* node, bool names, bool lens, bool, bool expand) {
    std::ostringstream ss;
    write_subtree(ss, T, node, names, lens, expand);
    return ss.str();
}
std::string get_newick_string(const Tree& T, bool names, bool lens, bool keep, bool expand) {
    return get_newick_string(T, T.root, names, lens, keep, expand);
}

void string_split(const std::string& s, char delim, std::vector<std::string>& words) {
    size_t a = 0, b;
    while ((b = s.find(delim, a)) != std::string::npos) {
        words.emplace_back(s.substr(a, b - a));
        a = b + 1;
    }
    if (a < s.size()) words.emplace_back(s.substr(a));
}
void string_split(const std::string& s, std::vector<std::string>& words) {
    std::istringstream ss(s);
    std::string w;
    while (ss >> w) words.push_back(w);
}

// Parser with the reference's observable behaviour (:415-508): split on ',', leaf name = characters before the
// first ':' or ')', internal nodes are (re)named node_1, node_2, ... in order of their '(' (names written after
// ')' are ignored), branch lengths are parsed but never printed back.
Tree create_tree_from_newick_string(const std::string& nwk) {
    Tree T;
    {   // one node per ',' piece and one per '(': size the name index once
        size_t pieces = 1;
        for (char c : nwk) pieces += (c == ',' || c == '(');
        T.reserve_nodes(pieces);
    }
    std::vector<Node*> stack;
    long depth = 0;
    std::string leaf;
    // pieces between ',' (string_split semantics: a trailing empty piece is dropped), scanned in place
    for (size_t a = 0; a < nwk.size();) {
        size_t b = nwk.find(',', a);
        const bool last = b == std::string::npos;
        if (last) b = nwk.size();
        size_t opens = 0, closes = 0;
        leaf.clear();
        bool stop = false;
        for (size_t i = a; i < b; i++) {
            const char c = nwk[i];
            if (c == '(') opens++;
            else if (c == ')') { closes++; stop = true; }
            else if (c == ':') stop = true;
            else if (!stop) leaf += c;
        }
        a = last ? nwk.size() : b + 1;
        for (size_t j = 0; j < opens; j++) {
            const std::string nid = T.new_internal_node_id();
            Node* nn = stack.empty() ? T.create_node(nid, -1.0f) : T.create_node(nid, stack.back(), -1.0f);
            stack.push_back(nn);
            depth++;
        }
        if (stack.empty()) {
            fprintf(stderr, "ERROR: incorrect Newick format!\n");
            exit(1);
        }
        T.create_node(leaf, stack.back(), -1.0f);
        for (size_t j = 0; j < closes; j++) {
            if (stack.empty()) { fprintf(stderr, "ERROR: incorrect Newick format!\n"); exit(1); }
            stack.pop_back();
            depth--;
        }
    }
    if (depth != 0) {
        fprintf(stderr, "ERROR: incorrect Newick format!\n");
        exit(1);
    }
    if (!T.root) fprintf(stderr, "WARNING: Tree found empty!\n");
    return T;
}
Tree create_tree_from_newick(const std::string& filename) {
    std::ifstream f(filename);
    if (!f) {
        fprintf(stderr, "ERROR: Could not open the tree file: %s!\n", filename.c_str());
        exit(1);
    }
    std::string line;
    std::getline(f, line);
    return create_tree_from_newick_string(line);
}

// ---------------------------------------------------------------- file helpers (plain or gzip through zlib)
static bool read_all(const std::string& filename, std::string& out) {
    gzFile f = gzopen(filename.c_str(), "rb");   // transparently reads uncompressed files too
    if (!f) return false;
    char buf[1 << 16];
    int n;
    while ((n = gzread(f, buf, sizeof buf)) > 0) out.append(buf, (size_t)n);
    gzclose(f);
    return n == 0;
}
// The bytes of a file: an uncompressed file is mapped (a 10 M-node protobuf is > 3 GB; no copy), a gzip one inflated.
struct FileBytes {
    const char* data = nullptr;
    size_t size = 0;
    std::string owned;
    void* map = nullptr;
    size_t map_len = 0;
    FileBytes() = default;
    FileBytes(const FileBytes&) = delete;
    FileBytes& operator=(const FileBytes&) = delete;
    ~FileBytes() { if (map) munmap(map, map_len); }
    bool open(const std::string& filename) {
        int fd = ::open(filename.c_str(), O_RDONLY);
        if (fd < 0) return false;
        unsigned char magic[2] = {0, 0};
        struct stat st;
        const bool plain = fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0 &&
                           !(pread(fd, magic, 2, 0) == 2 && magic[0] == 0x1f && magic[1] == 0x8b);
        if (plain) {
            void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
            if (m != MAP_FAILED) {
                map = m; map_len = (size_t)st.st_size;
                data = (const char*)m; size = map_len;
                ::close(fd);
                return true;
            }
        }
        ::close(fd);
        if (!read_all(filename, owned)) return false;
        data = owned.data(); size = owned.size();
        return true;
    }
};
static bool write_all(const std::string& filename, const std::string& data) {
    if (filename.find(".gz") != std::string::npos) {
        gzFile f = gzopen(filename.c_str(), "wb");
        if (!f) return false;
        size_t off = 0;
        while (off < data.size()) {
            int n = gzwrite(f, data.data() + off, (unsigned)std::min<size_t>(data.size() - off, 1u << 30));
            if (n <= 0) { gzclose(f); return false; }
            off += (size_t)n;
        }
        return gzclose(f) == Z_OK;
    }
    std::ofstream o(filename, std::ios::binary);
    if (!o) return false;
    o.write(data.data(), (std::streamsize)data.size());
    return (bool)o;
}

// ---------------------------------------------------------------- parsimony.proto wire codec
namespace pb {
struct Reader {
    const uint8_t* p; const uint8_t* end; bool ok = true;
    Reader(const void* d, size_t n) : p((const uint8_t*)d), end((const uint8_t*)d + n) {}
    bool more() const { return ok && p < end; }
    uint64_t varint() {
        uint64_t v = 0; int sh = 0;
        while (p < end) {
            uint8_t b = *p++;
            v |= (uint64_t)(b & 0x7f) << sh;
            if (!(b & 0x80)) return v;
            sh += 7;
            if (sh > 63) break;
        }
        ok = false; return 0;
    }
    Reader sub() {
        uint64_t n = varint();
        if (!ok || n > (uint64_t)(end - p)) { ok = false; return Reader(p, 0); }
        Reader r(p, (size_t)n); p += n; return r;
    }
    void skip(uint32_t wt) {
        switch (wt) {
            case 0: varint(); break;
            case 1: if (end - p >= 8) p += 8; else ok = false; break;
            case 2: sub(); break;
            case 5: if (end - p >= 4) p += 4; else ok = false; break;
            default: ok = false;
        }
    }
};
inline void put_varint(std::string& o, uint64_t v) {
    while (v >= 0x80) { o.push_back((char)(v | 0x80)); v >>= 7; }
    o.push_back((char)v);
}
inline void put_tag(std::string& o, uint32_t field, uint32_t wt) { put_varint(o, (field << 3) | wt); }
inline void put_bytes(std::string& o, uint32_t field, const std::string& s) {
    put_tag(o, field, 2); put_varint(o, s.size()); o += s;
}
inline void put_i32(std::string& o, uint32_t field, int32_t v) {   // proto3: default (0) is not written
    if (v == 0) return;
    put_tag(o, field, 0); put_varint(o, (uint64_t)(int64_t)v);
}
}  // namespace pb

Tree load_mutation_annotated_tree(const std::string& filename) {   // reference :522-612
    const bool timing = getenv("UB200_LOAD_TIMING") != nullptr;   // developer switch: phase times on stderr
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[pb load] %-24s %7.1f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    };
    FileBytes raw;
    if (!raw.open(filename)) {
        fprintf(stderr, "ERROR: Could not load the mutation-annotated tree object from file: %s!\n", filename.c_str());
        exit(1);
    }
    pb::Reader top(raw.data, raw.size);
    std::string newick;
    std::vector<pb::Reader> lists, conds, metas;
    while (top.more()) {
        const uint64_t tag = top.varint();
        const uint32_t field = (uint32_t)(tag >> 3), wt = (uint32_t)(tag & 7);
        if (wt == 2 && field >= 1 && field <= 4) {
            pb::Reader r = top.sub();
            if (field == 1) newick.assign((const char*)r.p, (size_t)(r.end - r.p));
            else if (field == 2) lists.push_back(r);
            else if (field == 3) conds.push_back(r);
            else metas.push_back(r);
        } else {
            top.skip(wt);
        }
    }
    if (!top.ok) {
        fprintf(stderr, "ERROR: %s is not a valid parsimony.proto message\n", filename.c_str());
        exit(1);
    }
    if (metas.empty()) fprintf(stderr, "WARNING: This pb does not include any metadata. Filling in default values\n");
    lap("top-level fields");
    Tree tree = create_tree_from_newick_string(newick);
    lap("newick -> nodes");
    auto dfs = tree.depth_first_expansion();
    lap("depth-first expansion");
    if (lists.size() < dfs.size()) {
        fprintf(stderr, "ERROR: protobuf holds %zu mutation lists for %zu nodes\n", lists.size(), dfs.size());
        exit(1);
    }
    for (size_t i = 0; i < dfs.size(); i++) {
        Node* node = dfs[i];
        if (i < metas.size()) {
            pb::Reader m = metas[i];
            while (m.more()) {
                const uint64_t tag = m.varint();
                if ((tag >> 3) == 1 && (tag & 7) == 2) {
                    pb::Reader s = m.sub();
                    node->clade_annotations.emplace_back((const char*)s.p, (size_t)(s.end - s.p));
                } else m.skip((uint32_t)(tag & 7));
            }
        }
        pb::Reader l = lists[i];
        {   // entries of the list (field 1, length-delimited): size the row once
            pb::Reader c = l;
            size_t entries = 0;
            while (c.more()) {
                const uint64_t tag = c.varint();
                entries += ((tag >> 3) == 1 && (tag & 7) == 2);
                c.skip((uint32_t)(tag & 7));
            }
            node->mutations.reserve(entries);
        }
        while (l.more()) {
            const uint64_t tag = l.varint();
            if (!((tag >> 3) == 1 && (tag & 7) == 2)) { l.skip((uint32_t)(tag & 7)); continue; }
            pb::Reader mr = l.sub();
            int32_t pos = 0, refn = 0, parn = 0;
            int8_t mut = 0;   // get_nuc_id(vector): the sum of 1 << code over the repeated field
            Mutation m;
            while (mr.more()) {
                const uint64_t t2 = mr.varint();
                const uint32_t f2 = (uint32_t)(t2 >> 3), w2 = (uint32_t)(t2 & 7);
                if (f2 == 1 && w2 == 0) pos = (int32_t)mr.varint();
                else if (f2 == 2 && w2 == 0) refn = (int32_t)mr.varint();
                else if (f2 == 3 && w2 == 0) parn = (int32_t)mr.varint();
                else if (f2 == 4 && w2 == 0) mut = (int8_t)(mut + (1 << (int8_t)mr.varint()));
                else if (f2 == 4 && w2 == 2) { pb::Reader pk = mr.sub(); while (pk.more()) mut = (int8_t)(mut + (1 << (int8_t)pk.varint())); }
                else if (f2 == 5 && w2 == 2) { pb::Reader s = mr.sub(); m.chrom.assign((const char*)s.p, (size_t)(s.end - s.p)); }
                else mr.skip(w2);
            }
            m.position = pos;
            if (!m.is_masked()) {
                m.ref_nuc = (int8_t)(1 << refn);
                m.par_nuc = (int8_t)(1 << parn);
                m.mut_nuc = mut;
                if (m.mut_nuc == m.par_nuc) continue;   // :580
            } else {
                m.ref_nuc = m.par_nuc = m.mut_nuc = 0;
            }
            // rows are stored sorted: append without the search when the entry belongs at the end
            if (node->mutations.empty() || node->mutations.back().position < m.position) node->mutations.push_back(std::move(m));
            else node->add_mutation(m);
        }
        if (!std::is_sorted(node->mutations.begin(), node->mutations.end())) {
            fprintf(stderr, "WARNING: Mutations not sorted!\n");
            std::sort(node->mutations.begin(), node->mutations.end());
        }
    }
    lap("mutation lists, metadata");
    for (auto c : conds) {
        std::string name;
        std::vector<std::string> members;
        while (c.more()) {
            const uint64_t tag = c.varint();
            if ((tag & 7) != 2) { c.skip((uint32_t)(tag & 7)); continue; }
            pb::Reader s = c.sub();
            if ((tag >> 3) == 1) name.assign((const char*)s.p, (size_t)(s.end - s.p));
            else if ((tag >> 3) == 2) members.emplace_back((const char*)s.p, (size_t)(s.end - s.p));
        }
        for (auto& m : members) tree.condensed_leaves.insert(m);
        tree.condensed_nodes[name] = members;
        tree.condensed_order.push_back(name);
    }
    return tree;
}

void save_mutation_annotated_tree(const Tree& tree, const std::string& filename) {   // reference :614-681
    std::string out;
    pb::put_bytes(out, 1, get_newick_string(tree, false, true, true));
    auto dfs = tree.depth_first_expansion();
    for (auto n : dfs) {
        std::string list;
        for (auto& m : n->mutations) {
            std::string mm;
            pb::put_i32(mm, 1, m.position);
            if (m.is_masked()) {
                pb::put_i32(mm, 2, -1);
                pb::put_i32(mm, 3, -1);
            } else {
                pb::put_i32(mm, 2, get_nt(m.ref_nuc));
                pb::put_i32(mm, 3, get_nt(m.par_nuc));
                std::string packed;
                for (auto b : get_nuc_vec_from_id(m.mut_nuc)) pb::put_varint(packed, (uint64_t)b);
                if (!packed.empty()) pb::put_bytes(mm, 4, packed);
            }
            if (!m.chrom.empty()) pb::put_bytes(mm, 5, m.chrom);
            pb::put_bytes(list, 1, mm);
        }
        pb::put_bytes(out, 2, list);
    }
    std::vector<std::string> order = tree.condensed_order;
    for (auto& kv : tree.condensed_nodes)
        if (std::find(order.begin(), order.end(), kv.first) == order.end()) order.push_back(kv.first);
    for (auto& name : order) {
        auto it = tree.condensed_nodes.find(name);
        if (it == tree.condensed_nodes.end()) continue;
        std::string c;
        pb::put_bytes(c, 1, name);
        for (auto& l : it->second) pb::put_bytes(c, 2, l);
        pb::put_bytes(out, 3, c);
    }
    for (auto n : dfs) {
        std::string meta;
        for (auto& a : n->clade_annotations) pb::put_bytes(meta, 1, a);
        pb::put_bytes(out, 4, meta);
    }
    if (!write_all(filename, out)) {
        fprintf(stderr, "ERROR: Could not write %s\n", filename.c_str());
        exit(1);
    }
}

// ---------------------------------------------------------------- flat loader / saver (flat_mat.hpp)
bool load_flat_mutation_annotated_tree(const std::string& filename, FlatTree& t, std::string& err) {
    const bool timing = getenv("UB200_LOAD_TIMING") != nullptr;   // developer switch: phase times on stderr
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[flat load] %-24s %7.1f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    };
    FileBytes raw;
    if (!raw.open(filename)) { err = "could not read " + filename; return false; }
    lap("read file");
    pb::Reader top(raw.data, raw.size);
    pb::Reader nwk(nullptr, 0);
    std::vector<pb::Reader> lists, conds, metas;
    while (top.more()) {
        const uint64_t tag = top.varint();
        const uint32_t field = (uint32_t)(tag >> 3), wt = (uint32_t)(tag & 7);
        if (wt == 2 && field >= 1 && field <= 4) {
            pb::Reader r = top.sub();
            if (field == 1) nwk = r;
            else if (field == 2) lists.push_back(r);
            else if (field == 3) conds.push_back(r);
            else metas.push_back(r);
        } else {
            top.skip(wt);
        }
    }
    if (!top.ok) { err = filename + " is not a valid parsimony.proto message"; return false; }
    lap("top-level fields");
    // ---- newick -> parent[] / names[]: the reference's tokeniser (split at ',', '(' opens an internal node, the
    // text before the first ':' or ')' names the leaf, ')' closes), nodes numbered as they are created
    t = FlatTree();
    {
        const char* p = (const char*)nwk.p;
        const char* end = (const char*)nwk.end;
        std::vector<int32_t> stack;
        size_t internal = 0;
        while (p < end) {
            const char* q = p;
            while (q < end && *q != ',') q++;
            size_t opens = 0, closes = 0;
            std::string leaf;
            bool stop = false;
            for (const char* c = p; c < q; c++) {
                if (*c == '(') opens++;
                else if (*c == ')') { closes++; stop = true; }
                else if (*c == ':') stop = true;
                else if (!stop) leaf += *c;
            }
            for (size_t j = 0; j < opens; j++) {
                t.parent.push_back(stack.empty() ? -1 : stack.back());
                t.names.push_back("node_" + std::to_string(++internal));
                stack.push_back((int32_t)t.parent.size() - 1);
            }
            if (stack.empty()) { err = "incorrect Newick format"; return false; }
            t.parent.push_back(stack.back());
            t.names.push_back(leaf);
            for (size_t j = 0; j < closes; j++) {
                if (stack.empty()) { err = "incorrect Newick format"; return false; }
                stack.pop_back();
            }
            p = q + 1;
        }
        if (!stack.empty()) { err = "incorrect Newick format"; return false; }
    }
    lap("newick");
    const size_t n = t.parent.size();
    if (lists.size() < n) { err = "protobuf holds " + std::to_string(lists.size()) + " mutation lists for " + std::to_string(n) + " nodes"; return false; }
    t.n_children.assign(n, 0);
    for (size_t i = 1; i < n; i++) t.n_children[t.parent[i]]++;
    // ---- mutation lists -> CSR rows.  The lists are independent length-delimited fields (a 10 M-node tree is > 3 GB of
    // varints): slices of the node range are parsed on host threads, twice -- pass 1 counts the entries every row keeps
    // (and checks the chromosome names), pass 2 writes them straight into their place, so nothing is buffered or copied.
    t.row_ptr.assign(n + 1, 0);
    {
        unsigned nt = std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 16u);
        if (raw.size < (1u << 20)) nt = 1;
        if (const char* e = getenv("UB200_HOST_THREADS")) nt = (unsigned)std::max(1, atoi(e));   // tests force the split
        struct Part {
            bool chrom_set = false, any_unnamed = false;   // any_unnamed: a mutation without a chromosome name
            std::string chrom, err;
        };
        std::vector<Part> part(nt);
        // slices of roughly equal bytes
        std::vector<size_t> cut(nt + 1, n);
        cut[0] = 0;
        {
            const uint8_t* base = n ? lists[0].p : nullptr;
            const size_t span = n ? (size_t)(lists[n - 1].end - base) : 0;
            size_t i = 0;
            for (unsigned c = 1; c < nt; c++) {
                while (i < n && (size_t)(lists[i].p - base) < span * c / nt) i++;
                cut[c] = i;
            }
        }
        // one node's list; dst == nullptr: count only.  Returns the entries kept, or (size_t)-1 after setting P.err.
        auto parse_list = [&](size_t i, Part& P, ub200_mutation* dst) -> size_t {
            pb::Reader l = lists[i];
            size_t kept = 0;
            while (l.more()) {
                const uint64_t tag = l.varint();
                if (!((tag >> 3) == 1 && (tag & 7) == 2)) { l.skip((uint32_t)(tag & 7)); continue; }
                pb::Reader mr = l.sub();
                int32_t pos = 0, refn = 0, parn = 0;
                int8_t mut = 0;
                bool bad_code = false;
                const char* chrom_p = nullptr;
                size_t chrom_n = 0;
                while (mr.more()) {
                    const uint64_t t2 = mr.varint();
                    const uint32_t f2 = (uint32_t)(t2 >> 3), w2 = (uint32_t)(t2 & 7);
                    if (f2 == 1 && w2 == 0) pos = (int32_t)mr.varint();
                    else if (f2 == 2 && w2 == 0) refn = (int32_t)mr.varint();
                    else if (f2 == 3 && w2 == 0) parn = (int32_t)mr.varint();
                    else if (f2 == 4 && w2 == 0) { const uint64_t v = mr.varint(); bad_code |= v > 3; mut |= (int8_t)(1 << (int)(v & 3)); }
                    else if (f2 == 4 && w2 == 2) {
                        pb::Reader pk = mr.sub();
                        while (pk.more()) { const uint64_t v = pk.varint(); bad_code |= v > 3; mut |= (int8_t)(1 << (int)(v & 3)); }
                    }
                    else if (f2 == 5 && w2 == 2) { pb::Reader s = mr.sub(); chrom_p = (const char*)s.p; chrom_n = (size_t)(s.end - s.p); }
                    else mr.skip(w2);
                }
                if (!l.ok || !mr.ok) { P.err = "malformed mutation list of node " + std::to_string(i); return (size_t)-1; }
                if (pos >= 0 && (bad_code || (uint32_t)refn > 3u || (uint32_t)parn > 3u)) {
                    P.err = "node " + std::to_string(i) + ": nucleotide code outside 0..3 at position " + std::to_string(pos);
                    return (size_t)-1;
                }
                if (!dst) {
                    if (!chrom_n) P.any_unnamed = true;
                    if (chrom_n || P.chrom_set) {
                        if (!P.chrom_set) { P.chrom.assign(chrom_p ? chrom_p : "", chrom_n); P.chrom_set = true; }
                        else if (P.chrom.size() != chrom_n || (chrom_n && memcmp(P.chrom.data(), chrom_p, chrom_n) != 0)) {
                            P.err = "the tree names more than one chromosome: not representable in the flat form";
                            return (size_t)-1;
                        }
                    }
                }
                ub200_mutation m;
                m.position = pos;
                m.is_missing = 0;
                if (pos >= 0) {
                    m.ref_nuc = (uint8_t)(1 << refn);
                    m.par_nuc = (uint8_t)(1 << parn);
                    m.mut_nuc = (uint8_t)mut;
                    if (m.mut_nuc == m.par_nuc) continue;   // :580: entries that change nothing are dropped
                } else {
                    m.ref_nuc = m.par_nuc = m.mut_nuc = 0;
                }
                if (dst) dst[kept] = m;
                kept++;
            }
            return kept;
        };
        auto run = [&](auto body) {
            if (nt == 1) { body(0u); return; }
            std::vector<std::thread> pool;
            for (unsigned c = 0; c < nt; c++) pool.emplace_back(body, c);
            for (auto& th : pool) th.join();
        };
        run([&](unsigned c) {
            Part P;   // filled locally, published at the end: neighbouring records would share cache lines
            for (size_t i = cut[c]; i < cut[c + 1]; i++) {
                const size_t k = parse_list(i, P, nullptr);
                if (k == (size_t)-1) break;
                t.row_ptr[i + 1] = k;   // row length; offsets follow below
            }
            part[c] = std::move(P);
        });
        lap("  count (threads)");
        bool chrom_set = false;
        for (auto& P : part) {   // first error in node order; one chromosome name across the slices
            if (!P.err.empty()) { err = P.err; return false; }
            // (as in one sequential pass: once a name has been seen, every later mutation must carry the same one)
            if (chrom_set && (P.any_unnamed || (P.chrom_set && t.chrom != P.chrom))) {
                err = "the tree names more than one chromosome: not representable in the flat form";
                return false;
            }
            if (P.chrom_set && !chrom_set) { t.chrom = P.chrom; chrom_set = true; }
        }
        for (size_t i = 0; i < n; i++) t.row_ptr[i + 1] += t.row_ptr[i];
        t.muts.resize(t.row_ptr[n]);
        lap("  offsets, resize");
        run([&](unsigned c) {
            Part P;
            auto by_pos = [](const ub200_mutation& x, const ub200_mutation& y) { return x.position < y.position; };
            for (size_t i = cut[c]; i < cut[c + 1]; i++) {
                ub200_mutation* row = t.muts.data() + t.row_ptr[i];
                const size_t k = parse_list(i, P, row);
                // rows are stored position-sorted (Node::add_mutation keeps them so; masked entries first)
                if (!std::is_sorted(row, row + k, by_pos)) std::stable_sort(row, row + k, by_pos);
            }
        });
        lap("  fill (threads)");
    }
    lap("mutation lists");
    t.have_metadata = !metas.empty();
    t.annotations.assign(n, {});
    for (size_t i = 0; i < n && i < metas.size(); i++) {
        pb::Reader m = metas[i];
        while (m.more()) {
            const uint64_t tag = m.varint();
            if ((tag >> 3) == 1 && (tag & 7) == 2) {
                pb::Reader s = m.sub();
                t.annotations[i].emplace_back((const char*)s.p, (size_t)(s.end - s.p));
            } else m.skip((uint32_t)(tag & 7));
        }
    }
    for (auto c : conds) {
        std::string name;
        std::vector<std::string> members;
        while (c.more()) {
            const uint64_t tag = c.varint();
            if ((tag & 7) != 2) { c.skip((uint32_t)(tag & 7)); continue; }
            pb::Reader s = c.sub();
            if ((tag >> 3) == 1) name.assign((const char*)s.p, (size_t)(s.end - s.p));
            else if ((tag >> 3) == 2) members.emplace_back((const char*)s.p, (size_t)(s.end - s.p));
        }
        t.condensed.emplace_back(std::move(name), std::move(members));
    }
    lap("metadata, condensed");
    return true;
}

bool save_flat_mutation_annotated_tree(const FlatTree& t, const std::string& filename, std::string& err) {
    const size_t n = t.parent.size();
    std::string out;
    // ---- newick: leaves by name, no internal names, branch length = number of mutations (as save_mutation_annotated_tree)
    {
        std::string nwk;
        auto len_of = [&](size_t i) {
            std::ostringstream ss;
            ss << ':' << static_cast<float>(t.row_ptr[i + 1] - t.row_ptr[i]);
            return ss.str();
        };
        // children lists in DFS order
        std::vector<uint32_t> first(n + 1, 0), kids(n ? n - 1 : 0);
        for (size_t i = 1; i < n; i++) first[t.parent[i] + 1]++;
        for (size_t i = 0; i < n; i++) first[i + 1] += first[i];
        { std::vector<uint32_t> fill(first.begin(), first.end() - 1); for (size_t i = 1; i < n; i++) kids[fill[t.parent[i]]++] = (uint32_t)i; }
        struct Frame { uint32_t node; uint32_t next; };
        std::vector<Frame> st;
        if (n) st.push_back({0, 0});
        while (!st.empty()) {
            Frame& f = st.back();
            const uint32_t u = f.node, nk = first[u + 1] - first[u];
            if (nk == 0) { nwk += t.names[u]; nwk += len_of(u); st.pop_back(); continue; }
            if (f.next == 0) nwk += '(';
            if (f.next < nk) {
                if (f.next) nwk += ',';
                const uint32_t c = kids[first[u] + f.next++];
                st.push_back({c, 0});
                continue;
            }
            nwk += ')';
            nwk += len_of(u);
            st.pop_back();
        }
        nwk += ';';
        pb::put_bytes(out, 1, nwk);
    }
    auto code = [](uint8_t one_hot) { return (int32_t)(31 - __builtin_clz((unsigned)one_hot)); };
    {   // the per-node mutation lists are independent fields: slices of the node range are serialised on host threads
        // into their own buffers (reused scratch strings, no allocation per mutation) and appended in order
        unsigned nt = std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 16u);
        if (t.muts.size() < (1u << 16)) nt = 1;
        if (const char* e = getenv("UB200_HOST_THREADS")) nt = (unsigned)std::max(1, atoi(e));   // tests force the split
        std::vector<size_t> cut(nt + 1, n);
        cut[0] = 0;
        for (unsigned c = 1; c < nt; c++)
            cut[c] = (size_t)(std::lower_bound(t.row_ptr.begin(), t.row_ptr.begin() + n, t.muts.size() * c / nt) - t.row_ptr.begin());
        for (unsigned c = 1; c <= nt; c++) cut[c] = std::max(cut[c], cut[c - 1]);
        std::vector<std::string> bufs(nt);
        auto body = [&](unsigned c) {
            std::string buf, list, mm, packed;
            buf.reserve((size_t)(t.row_ptr[cut[c + 1]] - t.row_ptr[cut[c]]) * 14 + (cut[c + 1] - cut[c]) * 3 + 64);
            for (size_t i = cut[c]; i < cut[c + 1]; i++) {
                list.clear();
                for (uint64_t k = t.row_ptr[i]; k < t.row_ptr[i + 1]; k++) {
                    const ub200_mutation& m = t.muts[k];
                    mm.clear();
                    pb::put_i32(mm, 1, m.position);
                    if (m.position < 0) {
                        pb::put_i32(mm, 2, -1);
                        pb::put_i32(mm, 3, -1);
                    } else {
                        pb::put_i32(mm, 2, code(m.ref_nuc));
                        pb::put_i32(mm, 3, code(m.par_nuc));
                        packed.clear();
                        for (int b = 0; b < 4; b++) if (m.mut_nuc & (1 << b)) pb::put_varint(packed, (uint64_t)b);
                        if (!packed.empty()) pb::put_bytes(mm, 4, packed);
                    }
                    if (!t.chrom.empty()) pb::put_bytes(mm, 5, t.chrom);
                    pb::put_bytes(list, 1, mm);
                }
                pb::put_bytes(buf, 2, list);
            }
            bufs[c].swap(buf);
        };
        if (nt == 1) body(0);
        else {
            std::vector<std::thread> pool;
            for (unsigned c = 0; c < nt; c++) pool.emplace_back(body, c);
            for (auto& th : pool) th.join();
        }
        size_t total = out.size();
        for (auto& b : bufs) total += b.size();
        out.reserve(total + 1024);
        for (auto& b : bufs) { out += b; std::string().swap(b); }
    }
    for (auto& c : t.condensed) {
        std::string cc;
        pb::put_bytes(cc, 1, c.first);
        for (auto& l : c.second) pb::put_bytes(cc, 2, l);
        pb::put_bytes(out, 3, cc);
    }
    for (size_t i = 0; i < n; i++) {
        std::string meta;
        if (i < t.annotations.size()) for (auto& a : t.annotations[i]) pb::put_bytes(meta, 1, a);
        pb::put_bytes(out, 4, meta);
    }
    if (!write_all(filename, out)) { err = "could not write " + filename; return false; }
    return true;
}

// ---------------------------------------------------------------- outputs / inputs around the placement
void get_sample_mutation_paths(Tree* T, const std::vector<std::string>& samples, const std::string& filename) {
    FILE* f = fopen(filename.c_str(), "w");   // reference :1991-2050
    if (!f) { fprintf(stderr, "ERROR: cannot write %s\n", filename.c_str()); exit(1); }
    for (auto& sample : samples) {
        Node* sn = T->get_node(sample);
        if (!sn) continue;   // not placed (thresholds)
        std::vector<std::string> segs;
        auto seg = [&](Node* n) {
            if (n->mutations.empty()) return;
            std::string s = n->identifier + ":";
            for (size_t k = 0; k < n->mutations.size(); k++)
                s += n->mutations[k].get_string() + (k + 1 < n->mutations.size() ? "," : " ");
            segs.push_back(s);
        };
        seg(sn);
        for (auto a : T->rsearch(sample)) seg(a);
        fprintf(f, "%s\t", sample.c_str());
        for (auto it = segs.rbegin(); it != segs.rend(); ++it) fputs(it->c_str(), f);
        fputc('\n', f);
    }
    fclose(f);
}

// MAT construction: one Fitch-Sankoff pass per VCF row over the BFS vector (reference mapper_body,
// src/usher_mapper.cpp:6-161, driven row by row from read_vcf :2099-2179).  Serial host code: this is the
// pre-processing stage of `usher -t tree.nh -v samples.vcf -o tree.pb`, not the placement hot path.
// ---- VCF text: (pointer, length) views of the mapped file instead of getline + split copies
namespace {
struct Tok { const char* p; size_t n; };
inline bool vcf_space(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f' || c == '\n'; }
// whitespace-separated words of [p, le), like `istringstream >> word`
inline void vcf_words(const char* p, const char* le, std::vector<Tok>& w) {
    w.clear();
    for (const char* q = p; q < le;) {
        while (q < le && vcf_space(*q)) q++;
        if (q == le) break;
        const char* a0 = q;
        while (q < le && !vcf_space(*q)) q++;
        w.push_back(Tok{a0, (size_t)(q - a0)});
    }
}
// pieces of a word between ',' (a trailing empty piece is dropped, like the reference's string_split)
inline void vcf_alleles(const Tok& t, std::vector<Tok>& out) {
    out.clear();
    const char* a0 = t.p;
    const char* const ae = t.p + t.n;
    for (const char* x = a0; x < ae; x++)
        if (*x == ',') { out.push_back(Tok{a0, (size_t)(x - a0)}); a0 = x + 1; }
    if (a0 < ae) out.push_back(Tok{a0, (size_t)(ae - a0)});
}
// std::stoi on a word that starts with [+-]digits
inline long vcf_int(const Tok& t, const char* what) {
    size_t i = 0;
    bool neg = false;
    if (i < t.n && (t.p[i] == '+' || t.p[i] == '-')) neg = t.p[i++] == '-';
    if (i >= t.n || !isdigit((unsigned char)t.p[i])) {
        fprintf(stderr, "ERROR! Incorrect VCF format: %s '%.*s' is not a number.\n", what, (int)t.n, t.p);
        exit(1);
    }
    long v = 0;
    for (; i < t.n && isdigit((unsigned char)t.p[i]); i++) v = std::min<long>(v * 10 + (t.p[i] - '0'), 1L << 40);
    return neg ? -v : v;
}
}  // namespace

static void build_mat_from_vcf(Tree* T, const std::string& vcf_filename, std::vector<Missing_Sample>& missing_samples) {
    std::vector<Node*> bfs = T->breadth_first_expansion();
    std::unordered_map<std::string, size_t> bfs_idx;
    for (size_t i = 0; i < bfs.size(); i++) bfs_idx[bfs[i]->identifier] = i;
    const int big = (int)bfs.size();
    std::vector<size_t> parent_idx(bfs.size(), (size_t)-1);
    std::vector<uint8_t> is_leaf(bfs.size());
    for (size_t i = 0; i < bfs.size(); i++) {
        if (bfs[i]->parent) parent_idx[i] = bfs_idx[bfs[i]->parent->identifier];
        is_leaf[i] = bfs[i]->is_leaf();
    }
    fprintf(stderr, "Loading VCF file.\n");
    FileBytes raw;
    if (!raw.open(vcf_filename)) {
        fprintf(stderr, "ERROR: Could not open the VCF file: %s!\n", vcf_filename.c_str());
        exit(1);
    }
    fprintf(stderr, "Computing parsimonious assignments for input variants.\n");
    bool header = false;
    size_t n_ids = 0;
    std::vector<long> col_node;      // BFS index of the column's sample, or -(1+k) for missing sample k
    // Sites are parsed first; the assignment itself runs on the GPU (ub200_fs_*, one CTA per site).  The serial
    // restatement further down only runs when UB200_FS_HOST=1 asks for it.
    struct Site { int pos; int8_t ref; std::string chrom; size_t v0, v1; };
    std::vector<Site> sites;
    std::vector<uint32_t> var_node;
    std::vector<uint8_t> var_nuc;
    std::vector<Tok> w, alleles;
    const char* const end = raw.data + raw.size;
    for (const char* p = raw.data; p < end;) {
        const char* le = (const char*)memchr(p, '\n', (size_t)(end - p));
        if (!le) le = end;
        vcf_words(p, le, w);
        p = le < end ? le + 1 : end;
        if (!header) {
            if (w.size() > 1 && w[1].n == 3 && memcmp(w[1].p, "POS", 3) == 0) {
                for (size_t j = 9; j < w.size(); j++) {
                    const std::string name(w[j].p, w[j].n);
                    n_ids++;
                    auto it = bfs_idx.find(name);
                    if (it == bfs_idx.end()) {
                        missing_samples.emplace_back(Missing_Sample(name));
                        col_node.push_back(-(long)missing_samples.size());
                    } else {
                        col_node.push_back((long)it->second);
                    }
                }
                header = true;
            }
            continue;
        }
        if (w.size() != 9 + n_ids) {
            fprintf(stderr, "ERROR! Incorrect VCF format.\n");
            exit(1);
        }
        Site st;
        st.pos = (int)vcf_int(w[1], "POS");
        st.ref = get_nuc_id(w[3].p[0]);
        st.chrom.assign(w[0].p, w[0].n);
        st.v0 = var_node.size();
        vcf_alleles(w[4], alleles);
        fprintf(stderr, "At variant site %i\n", st.pos);
        for (size_t c = 0; c < n_ids; c++) {
            const Tok& gt = w[9 + c];
            int8_t nuc;
            if (isdigit((unsigned char)gt.p[0])) {
                const long a = vcf_int(gt, "genotype");
                if (a <= 0) continue;
                if ((size_t)a > alleles.size()) {
                    fprintf(stderr, "ERROR! Incorrect VCF format: genotype %ld at position %d has no ALT allele.\n", a, st.pos);
                    exit(1);
                }
                const Tok& al = alleles[(size_t)a - 1];
                nuc = get_nuc_id(al.n ? al.p[0] : '\0');   // first character only, like the reference
            } else {
                nuc = 15;
            }
            if (col_node[c] >= 0) {
                var_node.push_back((uint32_t)col_node[c]);
                var_nuc.push_back((uint8_t)nuc);
            } else {
                Mutation m;
                m.chrom = st.chrom;
                m.position = st.pos;
                m.ref_nuc = st.ref;
                m.par_nuc = st.ref;   // the reference leaves par_nuc unset here (:65-82); scoring never reads it
                m.is_missing = (nuc == 15);
                m.mut_nuc = nuc;
                missing_samples[(size_t)(-col_node[c] - 1)].mutations.push_back(std::move(m));
            }
        }
        st.v1 = var_node.size();
        sites.push_back(st);
    }
    auto add = [&](const Site& st, size_t node, int par_state, int state) {
        Mutation m;
        m.chrom = st.chrom;
        m.position = st.pos;
        m.ref_nuc = st.ref;
        m.par_nuc = (int8_t)(1 << par_state);
        m.mut_nuc = (int8_t)(1 << state);
        bfs[node]->add_mutation(m);
    };
    // No silent CPU path: without a device the run stops, unless the serial host restatement below is asked for by name
    // (UB200_FS_HOST=1: a diagnostic, and what the CPU-only tests of the readers / condense / save around it use).
    const bool host_asked = getenv("UB200_FS_HOST") != nullptr;
    if (!host_asked && ub200_device_count() <= 0) {
        fprintf(stderr, "ERROR: no CUDA device: the per-site parsimony assignment of the create-MAT mode (-t with -v) runs on "
                        "the GPU.  (UB200_FS_HOST=1 runs the serial host restatement instead; diagnostic only.)\n");
        exit(1);
    }
    const bool on_gpu = !host_asked;
    if (on_gpu) {
        fprintf(stderr, "Fitch-Sankoff on the GPU: %zu sites x %zu nodes\n", sites.size(), bfs.size());
        std::vector<int32_t> par(bfs.size());
        for (size_t i = 0; i < bfs.size(); i++) par[i] = parent_idx[i] == (size_t)-1 ? -1 : (int32_t)parent_idx[i];
        ub200_fs_tree* F = nullptr;
        if (ub200_fs_tree_create((uint32_t)bfs.size(), par.data(), 0, &F) != UB200_OK) {
            fprintf(stderr, "ERROR: %s\n", ub200_last_error());
            exit(1);
        }
        const size_t kBatch = 4096;
        for (size_t s0 = 0; s0 < sites.size(); s0 += kBatch) {
            const size_t ns = std::min(kBatch, sites.size() - s0);
            std::vector<uint8_t> refc(ns);
            std::vector<uint64_t> ptr(ns + 1);
            for (size_t k = 0; k < ns; k++) { refc[k] = (uint8_t)get_nt(sites[s0 + k].ref); ptr[k] = sites[s0 + k].v0 - sites[s0].v0; }
            ptr[ns] = sites[s0 + ns - 1].v1 - sites[s0].v0;
            std::vector<uint32_t> o_site, o_node;
            std::vector<uint8_t> o_st;
            uint64_t cap = std::max<uint64_t>(4096, 2 * ptr[ns]), cnt = 0;
            for (;;) {
                o_site.resize(cap); o_node.resize(cap); o_st.resize(cap);
                const int rc = ub200_fs_sites(F, (uint32_t)ns, refc.data(), ptr.data(), var_node.data() + sites[s0].v0,
                                              var_nuc.data() + sites[s0].v0, cap, o_site.data(), o_node.data(), o_st.data(), &cnt);
                if (rc == UB200_E_CAPACITY) { cap = cnt + 16; continue; }
                if (rc != UB200_OK) { fprintf(stderr, "ERROR: %s\n", ub200_last_error()); exit(1); }
                break;
            }
            // mutations are added site by site, nodes in BFS order, as the reference's serial loop would
            std::vector<size_t> ord(cnt);
            for (size_t k = 0; k < cnt; k++) ord[k] = k;
            std::sort(ord.begin(), ord.end(), [&](size_t a, size_t b) {
                return o_site[a] != o_site[b] ? o_site[a] < o_site[b] : o_node[a] < o_node[b];
            });
            for (size_t k : ord) add(sites[s0 + o_site[k]], o_node[k], o_st[k] >> 4, o_st[k] & 15);
        }
        ub200_fs_tree_destroy(F);
        return;
    }
    std::vector<std::array<int, 4>> score(bfs.size());
    std::vector<int8_t> state(bfs.size());
    for (const Site& st : sites) {
        const int ref_nt = get_nt(st.ref);
        // leaves: reference allele free, everything else "impossible" (:33-44); internal nodes start at 0
        for (size_t i = 0; i < bfs.size(); i++) {
            for (int j = 0; j < 4; j++) score[i][j] = (is_leaf[i] && j != ref_nt) ? big : 0;
            state[i] = 0;
        }
        for (size_t k = st.v0; k < st.v1; k++)
            for (int j = 0; j < 4; j++) score[var_node[k]][j] = (var_nuc[k] & (1 << j)) ? 0 : big;
        // Sankoff forward pass: children before parents = reverse BFS (:86-111)
        for (size_t i = bfs.size(); i-- > 1;) {
            const size_t p = parent_idx[i];
            for (int j = 0; j < 4; j++) {
                int best = big + 1;
                for (int k = 0; k < 4; k++) best = std::min(best, score[i][k] + (k != j));
                score[p][j] += best;
            }
        }
        // backward pass: keep the parent's state unless another base is strictly cheaper (:114-156)
        for (size_t i = 0; i < bfs.size(); i++) {
            const int8_t par_state = parent_idx[i] == (size_t)-1 ? (int8_t)ref_nt : state[parent_idx[i]];
            int8_t s2 = par_state;
            int best = score[i][par_state];
            for (int j = 0; j < 4; j++) if (score[i][j] < best) { best = score[i][j]; s2 = (int8_t)j; }
            state[i] = s2;
            if (s2 != par_state) add(st, i, par_state, s2);
        }
    }
}

void read_vcf(Tree* T, const std::string& vcf_filename, std::vector<Missing_Sample>& missing_samples,
              bool create_new_mat) {
    if (create_new_mat) {
        build_mat_from_vcf(T, vcf_filename, missing_samples);
        return;
    }
    read_vcf_samples(vcf_filename, [T](const std::string& name) { return T->get_node(name) || T->condensed_leaves.count(name); },
                     missing_samples);
}

void read_vcf_samples(const std::string& vcf_filename, const std::function<bool(const std::string&)>& in_tree,
                      std::vector<Missing_Sample>& missing_samples) {
    fprintf(stderr, "Loading VCF file\n");   // reference :2180-2278
    const auto t_start = std::chrono::steady_clock::now();
    FileBytes raw;
    if (!raw.open(vcf_filename)) {
        fprintf(stderr, "ERROR: Could not open the VCF file: %s!\n", vcf_filename.c_str());
        exit(1);
    }
    // Same observable behaviour as the reference's getline + whitespace split + per-genotype Mutation objects, without
    // materialising them: a 10 000-sample VCF holds 3e8 genotype tokens, almost all of them "0".  Words are (pointer,
    // length) views of the mapped file; a Mutation is only built for a genotype that adds one.
    // ---- header: the first line whose second word is POS names the samples
    size_t n_ids = 0;
    std::vector<size_t> cols;
    const char* p = raw.data;
    const char* const end = raw.data + raw.size;
    {
        std::vector<Tok> w;
        bool header = false;
        while (p < end && !header) {
            const char* le = (const char*)memchr(p, '\n', (size_t)(end - p));
            if (!le) le = end;
            vcf_words(p, le, w);
            p = le < end ? le + 1 : end;
            if (w.size() > 1 && w[1].n == 3 && memcmp(w[1].p, "POS", 3) == 0) {
                for (size_t j = 9; j < w.size(); j++) {
                    const std::string name(w[j].p, w[j].n);
                    n_ids++;
                    if (!in_tree(name)) {
                        missing_samples.emplace_back(Missing_Sample(name));
                        cols.push_back(j);
                    } else {
                        fprintf(stderr, "WARNING: Ignoring sample %s as it is already in the tree.\n", name.c_str());
                    }
                }
                header = true;
            }
        }
    }
    // ---- rows: slices of whole lines on host threads.  The one thing that crosses rows is the reference's stale
    // genotype allele (below): a slice counts, per sample, the reference calls it sees before its first non-reference
    // genotype, and the merge adds them when the allele carried in from the slices before it is ambiguous.
    unsigned nt = std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 16u);
    if ((size_t)(end - p) < (1u << 20)) nt = 1;
    if (const char* e = getenv("UB200_HOST_THREADS")) nt = (unsigned)std::max(1, atoi(e));   // tests force the split
    std::vector<const char*> cut(nt + 1, end);
    cut[0] = p;
    for (unsigned c = 1; c < nt; c++) {
        const char* q = p + (size_t)(end - p) * c / nt;
        q = std::max(q, cut[c - 1]);
        const char* nl = q < end ? (const char*)memchr(q, '\n', (size_t)(end - q)) : nullptr;
        cut[c] = nl ? nl + 1 : end;
    }
    struct Slice {
        std::vector<std::vector<Mutation>> muts;   // per new sample, in row order
        std::vector<size_t> ambiguous;             // per new sample, without the leading reference calls
        bool seen = false;                         // a non-reference genotype was seen
        size_t lead_rows = 0, lead_cols = 0;       // whole rows / samples of the next row before it
        int8_t last = 0;                           // allele of the last non-reference genotype
        std::string err;
    };
    std::vector<Slice> slices(nt);
    auto scan = [&](unsigned c) {
        Slice S;
        S.muts.resize(cols.size());
        S.ambiguous.assign(cols.size(), 0);
        std::vector<Tok> w, alleles;
        size_t rows = 0;
        int8_t stale_mut_nuc = 0;   // see the genotype loop below
        for (const char* q = cut[c]; q < cut[c + 1] && S.err.empty();) {
            const char* le = (const char*)memchr(q, '\n', (size_t)(cut[c + 1] - q));
            if (!le) le = cut[c + 1];
            vcf_words(q, le, w);
            q = le < cut[c + 1] ? le + 1 : cut[c + 1];
            if (w.size() != 9 + n_ids) {
                S.err = "ERROR! Incorrect VCF format. Expected " + std::to_string(9 + n_ids) + " columns but got " +
                        std::to_string(w.size()) + ".";
                break;
            }
            if (cols.empty()) continue;   // (the reference only looks at a row's fields per new sample)
            // ALT alleles: split at ',' (a trailing empty piece is dropped, like the reference's string_split)
            vcf_alleles(w[4], alleles);
            const std::string chrom(w[0].p, w[0].n);
            const int position = (int)vcf_int(w[1], "POS");
            const int8_t ref_nuc = get_nuc_id(w[3].p[0]);
            for (size_t k = 0; k < cols.size(); k++) {
                const Tok& gt = w[cols[k]];
                // The reference leaves Mutation::mut_nuc uninitialised here (src/mutation_annotated_tree.cpp:2246) and
                // tests it for ambiguity even when the genotype is 0 (:2271); the object reuses the stack slot of the
                // previous iteration, so num_ambiguous (the -A sort key) counts the PREVIOUS genotype's allele for
                // reference calls.  Reproduced, since the sample order decides the final tree.
                int8_t mut_nuc = stale_mut_nuc;
                bool add = false, is_missing = false;
                if (isdigit((unsigned char)gt.p[0])) {
                    const long a = vcf_int(gt, "genotype");
                    if (a > 0) {
                        if ((size_t)a > alleles.size()) {
                            S.err = "ERROR! Incorrect VCF format: genotype " + std::to_string(a) + " at position " +
                                    std::to_string(position) + " has no ALT allele.";
                            break;
                        }
                        const Tok& al = alleles[(size_t)a - 1];
                        const char c0 = al.n ? al.p[0] : '\0';   // first character only, like the reference
                        mut_nuc = get_nuc_id(c0);
                        is_missing = (c0 == 'N') || mut_nuc == 15;
                        add = true;
                    }
                } else {
                    is_missing = true;
                    mut_nuc = 15;
                    add = true;
                }
                if (add) {
                    Mutation m;
                    m.chrom = chrom;
                    m.position = position;
                    m.ref_nuc = ref_nuc;
                    m.par_nuc = ref_nuc;
                    m.mut_nuc = mut_nuc;
                    m.is_missing = is_missing;
                    S.muts[k].push_back(std::move(m));
                    if (!S.seen) { S.seen = true; S.lead_rows = rows; S.lead_cols = k; }
                    S.last = mut_nuc;
                }
                // (reference calls before the slice's first non-reference genotype carry an allele from an earlier
                // slice: they are counted at the merge)
                if (S.seen && (mut_nuc & (mut_nuc - 1))) S.ambiguous[k]++;
                stale_mut_nuc = mut_nuc;
            }
            rows++;
        }
        if (!S.seen) { S.lead_rows = rows; S.lead_cols = 0; }
        slices[c] = std::move(S);
    };
    if (nt == 1) scan(0);
    else {
        std::vector<std::thread> pool;
        for (unsigned c = 0; c < nt; c++) pool.emplace_back(scan, c);
        for (auto& th : pool) th.join();
    }
    int8_t carried = 0;   // the allele the reference's reused Mutation object holds when a slice starts
    for (auto& S : slices) {
        if (!S.err.empty()) { fprintf(stderr, "%s\n", S.err.c_str()); exit(1); }
        const bool carried_ambiguous = (carried & (carried - 1)) != 0;
        for (size_t k = 0; k < cols.size(); k++) {
            if (carried_ambiguous) missing_samples[k].num_ambiguous += S.lead_rows + ((S.seen && k < S.lead_cols) ? 1 : 0);
            missing_samples[k].num_ambiguous += S.ambiguous[k];
            auto& dst = missing_samples[k].mutations;
            if (dst.empty()) dst = std::move(S.muts[k]);
            else for (auto& m : S.muts[k]) dst.push_back(std::move(m));
        }
        if (S.seen) carried = S.last;
    }
    if (getenv("UB200_LOAD_TIMING"))
        fprintf(stderr, "[vcf] %zu new samples, %.1f MB in %.1f ms\n", missing_samples.size(), raw.size / 1e6,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count());
}

}  // namespace Mutation_Annotated_Tree
