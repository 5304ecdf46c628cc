// Point sets of the host line tracer with the ITERATION ORDER of the reference's container.
//
// The reference keeps the points near a line in std::unordered_set<Point, PointHash> (include/semi_global_align.h:79-91)
// and sums doubles while iterating over it (scoreLineSegment, scorePointSet: semi_global_align.cpp:739-803), so the order
// in which the set hands out its elements is part of the result.  That order is a property of libstdc++'s _Hashtable:
//   * one singly linked list of all nodes; a bucket stores the node BEFORE its first node
//   * a new node goes to the front of its bucket, or, when the bucket is empty, to the front of the whole list
//     (_M_insert_bucket_begin, bits/hashtable.h)
//   * a rehash walks the list in its current order and rebuilds it with the same two rules (_M_rehash_aux, unique keys)
//   * bucket counts come from _Prime_rehash_policy (load factor 1): _M_need_rehash before every insertion of a new key
//   * range insertion / the range constructor are loops of single insertions (unique keys, GCC 13); a copy keeps
//     bucket count and list order
// PointSet below restates exactly that on two flat arrays (nodes with a next index, buckets with a "node before" index)
// and asks the very policy object of the standard library for the bucket counts, so the order is the reference's by
// construction; tests/cpp/test_pointset.cpp checks it element by element against std::unordered_set on random
// insertion sequences.  What changes is the cost: no node allocation, no pointer chasing through the heap, 32-bit
// remainders (coordinates are non-negative: hash codes fit 32 bits).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <functional>
#include <unordered_set>
#include <vector>

namespace ub200 {
namespace seed {

struct Point {
    int x, y;
    Point() : x(0), y(0) {}
    Point(int px, int py) : x(px), y(py) {}
    bool operator==(const Point& o) const { return x == o.x && y == o.y; }
    bool operator<(const Point& o) const { return x == o.x ? y < o.y : x < o.x; }
};
struct PointHash {  // include/semi_global_align.h:79-86
    size_t operator()(const Point& p) const { return (std::hash<int>()(p.x) ^ (std::hash<int>()(p.y) << 1)) >> 1; }
};

// Per-thread bump arena: the tracer builds and drops thousands of small sets per range; it is rewound when a range
// starts (a thread seeds one range at a time, and no set outlives its range).
struct Arena {
    std::vector<char*> blocks;
    size_t cur = 0, used = 0;
    static const size_t BLOCK = 1 << 20;
    ~Arena() { for (char* b : blocks) free(b); }
    void rewind() { cur = 0; used = 0; }
    static size_t rounded(size_t n) { return (n + 15) & ~(size_t)15; }
    void* alloc(size_t n) {
        n = rounded(n);
        if (n > BLOCK) return malloc(n);   // (not from the arena: released by deallocate)
        if (blocks.empty()) blocks.push_back((char*)malloc(BLOCK));
        if (used + n > BLOCK) {
            used = 0;
            if (++cur == blocks.size()) blocks.push_back((char*)malloc(BLOCK));
        }
        void* p = blocks[cur] + used;
        used += n;
        return p;
    }
    static Arena& mine() { static thread_local Arena a; return a; }
};
template <typename T>
struct ArenaAlloc {
    typedef T value_type;
    ArenaAlloc() {}
    template <typename U> ArenaAlloc(const ArenaAlloc<U>&) {}
    T* allocate(size_t n) { return (T*)Arena::mine().alloc(n * sizeof(T)); }
    void deallocate(T* p, size_t n) { if (Arena::rounded(n * sizeof(T)) > Arena::BLOCK) free(p); }
    template <typename U> bool operator==(const ArenaAlloc<U>&) const { return true; }
    template <typename U> bool operator!=(const ArenaAlloc<U>&) const { return false; }
};

class PointSet {
    struct Node { Point p; uint32_t code; int32_t next; };
    static const int32_t NIL = -1;           // end of list / empty bucket
    static const int32_t BEFORE_BEGIN = -2;  // "the node before the first node of the list"

public:
    class const_iterator {
    public:
        const_iterator(const PointSet* s, int32_t n) : s_(s), n_(n) {}
        const Point& operator*() const { return s_->nodes_[(size_t)n_].p; }
        const Point* operator->() const { return &s_->nodes_[(size_t)n_].p; }
        const_iterator& operator++() { n_ = s_->nodes_[(size_t)n_].next; return *this; }
        bool operator!=(const const_iterator& o) const { return n_ != o.n_; }
        bool operator==(const const_iterator& o) const { return n_ == o.n_; }
        typedef std::forward_iterator_tag iterator_category;
        typedef Point value_type;
        typedef std::ptrdiff_t difference_type;
        typedef const Point* pointer;
        typedef const Point& reference;
    private:
        const PointSet* s_;
        int32_t n_;
    };

    PointSet() : buckets_(1, NIL) {}
    template <typename It>
    PointSet(It first, It last) : buckets_(1, NIL) { insert(first, last); }

    size_t size() const { return nodes_.size(); }
    bool empty() const { return nodes_.empty(); }
    const_iterator begin() const { return const_iterator(this, head_); }
    const_iterator end() const { return const_iterator(this, NIL); }

    bool contains(const Point& p) const {
        const uint32_t code = codeOf(p);
        return findIn(code % (uint32_t)buckets_.size(), code, p) != NIL;
    }

    template <typename It>
    void insert(It first, It last) {
        for (; first != last; ++first) insert(*first);
    }

    // _M_insert_unique + _M_insert_unique_node (bits/hashtable.h)
    bool insert(const Point& p) {
        const uint32_t code = codeOf(p);
        uint32_t bkt = code % (uint32_t)buckets_.size();
        if (findIn(bkt, code, p) != NIL) return false;
        // (_M_need_rehash answers "no" without touching its state while the element count stays within _M_next_resize:
        // that first test of it is inlined here, everything else is asked from the library's own policy object)
        if (nodes_.size() + 1 > policy_._M_next_resize) {
            const std::pair<bool, std::size_t> grow = policy_._M_need_rehash(buckets_.size(), nodes_.size(), 1);
            if (grow.first) {
                rehash(grow.second);
                bkt = code % (uint32_t)buckets_.size();
            }
        }
        const int32_t id = (int32_t)nodes_.size();
        nodes_.push_back(Node{p, code, NIL});
        Node& nd = nodes_.back();
        const int32_t before = buckets_[bkt];
        if (before != NIL) {              // bucket not empty: behind its "node before"
            nd.next = nextOf(before);
            setNext(before, id);
        } else {                          // empty bucket: front of the whole list
            nd.next = head_;
            head_ = id;
            if (nd.next != NIL) buckets_[nodes_[(size_t)nd.next].code % (uint32_t)buckets_.size()] = id;
            buckets_[bkt] = BEFORE_BEGIN;
        }
        return true;
    }

private:
    // hash codes of points with non-negative coordinates fit 32 bits; anything else would need the 64-bit remainder
    static uint32_t codeOf(const Point& p) {
        const size_t h = PointHash()(p);
        if (h > 0xffffffffull) abort();
        return (uint32_t)h;
    }
    int32_t nextOf(int32_t n) const { return n == BEFORE_BEGIN ? head_ : nodes_[(size_t)n].next; }
    void setNext(int32_t n, int32_t v) { if (n == BEFORE_BEGIN) head_ = v; else nodes_[(size_t)n].next = v; }

    // _M_find_before_node: walks the nodes of one bucket (they are consecutive in the list)
    int32_t findIn(uint32_t bkt, uint32_t code, const Point& p) const {
        const int32_t before = buckets_[bkt];
        if (before == NIL) return NIL;
        const uint32_t nb = (uint32_t)buckets_.size();
        for (int32_t n = nextOf(before); n != NIL; n = nodes_[(size_t)n].next) {
            const Node& nd = nodes_[(size_t)n];
            if (nd.code == code && nd.p == p) return n;
            const int32_t nx = nd.next;
            if (nx == NIL || nodes_[(size_t)nx].code % nb != bkt) break;
        }
        return NIL;
    }

    // _M_rehash_aux(n, unique keys)
    void rehash(std::size_t n) {
        std::vector<int32_t, ArenaAlloc<int32_t> > nb(n, NIL);
        int32_t p = head_;
        head_ = NIL;
        uint32_t beginBkt = 0;
        while (p != NIL) {
            const int32_t next = nodes_[(size_t)p].next;
            const uint32_t bkt = nodes_[(size_t)p].code % (uint32_t)n;
            if (nb[bkt] == NIL) {
                nodes_[(size_t)p].next = head_;
                head_ = p;
                nb[bkt] = BEFORE_BEGIN;
                if (nodes_[(size_t)p].next != NIL) nb[beginBkt] = p;
                beginBkt = bkt;
            } else {
                const int32_t before = nb[bkt];
                if (before == BEFORE_BEGIN) { nodes_[(size_t)p].next = head_; head_ = p; }
                else { nodes_[(size_t)p].next = nodes_[(size_t)before].next; nodes_[(size_t)before].next = p; }
            }
            p = next;
        }
        buckets_.swap(nb);
    }

    std::vector<Node, ArenaAlloc<Node> > nodes_;
    std::vector<int32_t, ArenaAlloc<int32_t> > buckets_;
    int32_t head_ = NIL;
    std::__detail::_Prime_rehash_policy policy_;
};

typedef std::vector<Point> PointVector;

}  // namespace seed
}  // namespace ub200
