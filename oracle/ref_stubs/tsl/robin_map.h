// Stand-in for tsl::robin_map (absent here): alias of std::unordered_map.
// Only used to compile the reference's own sources into oracle/_ref.
#pragma once
#include <unordered_map>
#include <functional>
namespace tsl {
    template<typename K, typename V, typename H = std::hash<K>, typename E = std::equal_to<K>>
    using robin_map = std::unordered_map<K, V, H, E>;
}
