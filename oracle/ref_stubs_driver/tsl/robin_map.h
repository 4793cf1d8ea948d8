// Stand-in for tsl::robin_map (absent here) for the ref_driver build (oracle/Makefile): std::unordered_map plus the one extension the
// reference's ILP_input.cpp uses, a mutable iterator with value().  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <functional>
#include <unordered_map>
namespace tsl {
    template<typename K, typename V, typename H = std::hash<K>, typename E = std::equal_to<K>>
    class robin_map : public std::unordered_map<K, V, H, E> {
        using base = std::unordered_map<K, V, H, E>;
    public:
        using base::base;
        struct iterator : base::iterator {
            iterator() {}
            iterator(typename base::iterator it) : base::iterator(it) {}
            V& value() const { return (**this).second; }
            const K& key() const { return (**this).first; }
        };
        iterator find(const K& k) { return iterator(base::find(k)); }
        typename base::const_iterator find(const K& k) const { return base::find(k); }
        iterator begin() { return iterator(base::begin()); }
        iterator end() { return iterator(base::end()); }
        typename base::const_iterator begin() const { return base::begin(); }
        typename base::const_iterator end() const { return base::end(); }
    };
}
