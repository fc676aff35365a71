#pragma once
// TEST INFRASTRUCTURE ONLY (oracle build).  Minimal stand-in for the slice of the oneTBB API that the
// reference's usher_mapper.cpp / usher_common.cpp / mutation_annotated_tree.cpp touch, so that those
// files compile UNMODIFIED from /root/reference/src without oneTBB installed.  Scheduling only: no
// arithmetic lives here.  parallel_for = static contiguous chunks over std::thread.
#include <algorithm>
#include <functional>
#include <mutex>
#include <shared_mutex>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace tbb {

extern int oracle_threads;   // defined in oracle/ref_driver.cpp

struct mutex {
    std::mutex m_;
    void lock() { m_.lock(); }
    void unlock() { m_.unlock(); }
    bool try_lock() { return m_.try_lock(); }
};

struct rw_mutex {
    std::shared_mutex m_;
    struct scoped_lock {
        rw_mutex& owner_;
        bool exclusive_;
        scoped_lock(rw_mutex& o, bool write = true) : owner_(o), exclusive_(write) {
            if (exclusive_) owner_.m_.lock(); else owner_.m_.lock_shared();
        }
        ~scoped_lock() {
            if (exclusive_) owner_.m_.unlock(); else owner_.m_.unlock_shared();
        }
    };
};

template <class K, class V> using concurrent_unordered_map = std::unordered_map<K, V>;
template <class K> using concurrent_unordered_set = std::unordered_set<K>;

struct affinity_partitioner {};

template <class T> struct blocked_range {
    T lo_, hi_;
    size_t grain_;
    blocked_range(T lo, T hi, size_t grain = 1) : lo_(lo), hi_(hi), grain_(grain) {}
    T begin() const { return lo_; }
    T end() const { return hi_; }
    size_t size() const { return (size_t)(hi_ - lo_); }
};

template <class Range, class Body> void parallel_for(const Range& r, const Body& body) {
    const size_t n = r.size();
    const int nt = oracle_threads;
    if (nt <= 1 || n < 64) { body(r); return; }
    const size_t per = (n + nt - 1) / nt;
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; ++t) {
        auto lo = r.begin() + t * per;
        if ((size_t)(lo - r.begin()) >= n) break;
        auto hi = std::min(r.end(), (decltype(lo))(lo + per));
        pool.emplace_back([&body, lo, hi] { body(Range(lo, hi)); });
    }
    for (auto& th : pool) th.join();
}
template <class Range, class Body, class Part> void parallel_for(const Range& r, const Body& body, Part&) {
    parallel_for(r, body);
}
template <class It, class Cmp> void parallel_sort(It a, It b, Cmp c) { std::sort(a, b, c); }
template <class It> void parallel_sort(It a, It b) { std::sort(a, b); }

struct global_control {
    enum parameter { max_allowed_parallelism, thread_stack_size };
    global_control(parameter p, size_t v) { if (p == max_allowed_parallelism) oracle_threads = (int)v; }
};
namespace this_task_arena { inline int max_concurrency() { return (int)std::thread::hardware_concurrency(); } }
namespace info { inline int default_concurrency() { return (int)std::thread::hardware_concurrency(); } }

struct flow_control { bool stopped_ = false; void stop() { stopped_ = true; } };

namespace flow {
enum { unlimited = 0, serial = 1 };
struct graph {
    std::function<void()> pump_;
    void wait_for_all() { if (pump_) { auto p = pump_; pump_ = nullptr; p(); } }
};
template <class In, class Out> struct function_node {
    std::function<Out(In)> body_;
    template <class B> function_node(graph&, size_t, B b) : body_(b) {}
    bool try_put(const In& v) { body_(v); return true; }
};
template <class Out> struct input_node {
    graph& g_;
    std::function<Out(flow_control&)> body_;
    std::function<void(Out)> sink_;
    template <class B> input_node(graph& g, B b) : g_(g), body_(b) {}
    void activate() {
        g_.pump_ = [this] {
            flow_control fc;
            for (;;) { Out o = body_(fc); if (fc.stopped_) break; if (sink_) sink_(o); }
        };
    }
};
template <class Out, class Node> void make_edge(input_node<Out>& src, Node& dst) {
    Node* d = &dst;
    src.sink_ = [d](Out o) { d->body_(o); };
}
}  // namespace flow
}  // namespace tbb
