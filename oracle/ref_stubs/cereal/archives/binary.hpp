// oracle/ref_stubs/cereal/archives/binary.hpp -- minimal stand-in for cereal's binary archives (cereal is not in this image).
// TEST INFRASTRUCTURE ONLY: enough interface for the reference's bdd_cuda_base.h / .cu (save / load templates,
// include/bdd_solver/bdd_cuda_base.h:14-29, src/bdd_solver/bdd_cuda_base.cu:1486-1544) to compile; a real byte-stream archive so that
// a save / load round trip works, but not cereal's file format.
#pragma once
#include <cstddef>
#include <cstring>
#include <istream>
#include <ostream>
#include <type_traits>
#include <utility>
#include <vector>

namespace cereal {

namespace stub_detail {
template<typename T, typename A, typename = void> struct has_member_save : std::false_type {};
template<typename T, typename A> struct has_member_save<T, A, decltype(std::declval<const T&>().save(std::declval<A&>()), void())> : std::true_type {};
template<typename T, typename A, typename = void> struct has_member_load : std::false_type {};
template<typename T, typename A> struct has_member_load<T, A, decltype(std::declval<T&>().load(std::declval<A&>()), void())> : std::true_type {};
}

class BinaryOutputArchive {
public:
    explicit BinaryOutputArchive(std::ostream& s) : s_(s) {}
    template<typename... T> void operator()(const T&... v) { (put(v), ...); }
    template<typename T> BinaryOutputArchive& operator<<(const T& v) { put(v); return *this; }
private:
    template<typename T> typename std::enable_if<std::is_arithmetic<T>::value>::type put(const T& v) { s_.write(reinterpret_cast<const char*>(&v), sizeof(T)); }
    template<typename T> void put(const std::vector<T>& v)
    {
        const std::size_t n = v.size();
        put(n);
        for(const T& x : v) put(x);
    }
    // classes serialise through a member save(archive) when they have one (cereal's member form), else through a free save(archive, object)
    template<typename T> typename std::enable_if<!std::is_arithmetic<T>::value && stub_detail::has_member_save<T, BinaryOutputArchive>::value>::type put(const T& v) { v.save(*this); }
    template<typename T> typename std::enable_if<!std::is_arithmetic<T>::value && !stub_detail::has_member_save<T, BinaryOutputArchive>::value>::type put(const T& v) { save(*this, v); }
    std::ostream& s_;
};

class BinaryInputArchive {
public:
    explicit BinaryInputArchive(std::istream& s) : s_(s) {}
    template<typename... T> void operator()(T&... v) { (get(v), ...); }
    template<typename T> BinaryInputArchive& operator>>(T& v) { get(v); return *this; }
private:
    template<typename T> typename std::enable_if<std::is_arithmetic<T>::value>::type get(T& v) { s_.read(reinterpret_cast<char*>(&v), sizeof(T)); }
    template<typename T> void get(std::vector<T>& v)
    {
        std::size_t n = 0;
        get(n);
        v.resize(n);
        for(T& x : v) get(x);
    }
    template<typename T> typename std::enable_if<!std::is_arithmetic<T>::value && stub_detail::has_member_load<T, BinaryInputArchive>::value>::type get(T& v) { v.load(*this); }
    template<typename T> typename std::enable_if<!std::is_arithmetic<T>::value && !stub_detail::has_member_load<T, BinaryInputArchive>::value>::type get(T& v) { load(*this, v); }
    std::istream& s_;
};

} // namespace cereal
