// Stand-in for tsl::robin_set (absent here): alias of std::unordered_set.
#pragma once
#include <unordered_set>
#include <functional>
namespace tsl {
    template<typename K, typename H = std::hash<K>, typename E = std::equal_to<K>>
    using robin_set = std::unordered_set<K, H, E>;
}
