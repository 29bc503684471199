// TEST INFRASTRUCTURE ONLY -- the product path (recommendersystems_b200/, librwr_b200.so) never includes this file.
//
// The handful of .NET base-class-library types the reference's RWR path uses (Recommenders/RWRBased/Graph.cs, Model.cs,
// Recommender.cs need nothing else: Recommenders.csproj:69-83), written for the transliterated reference that
// oracle/cs2cpp.py emits into oracle/_ref/.  Semantics follow the BCL where the reference's code can observe them:
//   * Dictionary / List / arrays are REFERENCE types: copying the variable shares the object (Graph keeps the caller's
//     `nodes` and `edges` and reads them at buildGraph() time); a T[] variable may be null;
//   * `new T[n]` zero-initialises; structs have an implicit parameterless constructor (cs2cpp adds `S() = default;`);
//   * Dictionary's indexer getter throws KeyNotFoundException, Add throws ArgumentException on an existing key, an array
//     index outside [0, Length) throws IndexOutOfRangeException, a null array throws NullReferenceException;
//   * double.CompareTo orders NaN below everything and equal to itself; List.Sort(Comparison) sorts by the sign of the
//     comparison (the BCL's introsort is unstable, which cannot show here: the reference's comparison is a total order on
//     distinct keys, Recommender.cs:34-38).
// Arithmetic is untouched: the reference's expressions are compiled as they stand, in IEEE double (-ffp-contract=off).
#pragma once

#include <algorithm>
#include <cstdint>
#include <limits>
#include <memory>
#include <stdexcept>
#include <unordered_map>
#include <vector>

namespace bcl {

struct KeyNotFoundException : std::runtime_error { KeyNotFoundException() : std::runtime_error("KeyNotFoundException") {} };
struct ArgumentException : std::runtime_error { ArgumentException() : std::runtime_error("ArgumentException") {} };
struct IndexOutOfRangeException : std::runtime_error { IndexOutOfRangeException() : std::runtime_error("IndexOutOfRangeException") {} };
struct NullReferenceException : std::runtime_error { NullReferenceException() : std::runtime_error("NullReferenceException") {} };

// T[]
template <typename T>
class Array {
    std::shared_ptr<std::vector<T>> p_;
    std::vector<T>& v() const {
        if (!p_) throw NullReferenceException();
        return *p_;
    }

public:
    Array() {}
    Array(std::nullptr_t) {}
    explicit Array(long long n) : p_(std::make_shared<std::vector<T>>((size_t)(n < 0 ? throw ArgumentException() : n))) {}
    int Length() const { return (int)v().size(); }
    T& operator[](long long i) const {
        std::vector<T>& a = v();
        if (i < 0 || (size_t)i >= a.size()) throw IndexOutOfRangeException();
        return a[(size_t)i];
    }
    bool operator==(std::nullptr_t) const { return !p_; }
    bool operator!=(std::nullptr_t) const { return (bool)p_; }
    typename std::vector<T>::iterator begin() const { return v().begin(); }
    typename std::vector<T>::iterator end() const { return v().end(); }
};

// System.Collections.Generic.List<T>
template <typename T>
class List {
    std::shared_ptr<std::vector<T>> p_ = std::make_shared<std::vector<T>>();

public:
    List() {}
    void Add(const T& x) { p_->push_back(x); }
    bool Contains(const T& x) const { return std::find(p_->begin(), p_->end(), x) != p_->end(); }
    int Count() const { return (int)p_->size(); }
    T& operator[](long long i) const {
        if (i < 0 || (size_t)i >= p_->size()) throw IndexOutOfRangeException();      // ArgumentOutOfRangeException in the BCL
        return (*p_)[(size_t)i];
    }
    template <typename Cmp>
    void Sort(Cmp comparison) {
        std::sort(p_->begin(), p_->end(), [&](const T& a, const T& b) { return comparison(a, b) < 0; });
    }
    void Sort() { std::sort(p_->begin(), p_->end()); }
    typename std::vector<T>::iterator begin() const { return p_->begin(); }
    typename std::vector<T>::iterator end() const { return p_->end(); }
};

// System.Collections.Generic.Dictionary<K, V>
template <typename K, typename V>
class Dictionary {
    std::shared_ptr<std::unordered_map<K, V>> p_ = std::make_shared<std::unordered_map<K, V>>();

public:
    Dictionary() {}
    int Count() const { return (int)p_->size(); }
    bool ContainsKey(const K& k) const { return p_->find(k) != p_->end(); }
    void Add(const K& k, const V& v) {
        if (!p_->emplace(k, v).second) throw ArgumentException();
    }
    V& operator[](const K& k) const {          // the getter (the reference's path never assigns through the indexer)
        auto it = p_->find(k);
        if (it == p_->end()) throw KeyNotFoundException();
        return it->second;
    }
};

template <typename K, typename V>
struct KeyValuePair {
    K Key;
    V Value;
    KeyValuePair() : Key(), Value() {}
    KeyValuePair(const K& k, const V& v) : Key(k), Value(v) {}
};

// System.Double.CompareTo / System.Int64.CompareTo / System.Int32.CompareTo
inline int CompareTo(double a, double b) {
    if (a < b) return -1;
    if (a > b) return 1;
    if (a == b) return 0;
    if (a != a) return (b != b) ? 0 : -1;       // at least one NaN: NaN sorts below every number
    return 1;
}
inline int CompareTo(long long a, long long b) { return a < b ? -1 : (a > b ? 1 : 0); }
inline int CompareTo(int a, int b) { return a < b ? -1 : (a > b ? 1 : 0); }

}  // namespace bcl
