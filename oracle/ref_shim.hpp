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
//   * Dictionary.Keys / .Values and HashSet enumerate in insertion order (what the BCL does while nothing is removed; the
//     reference never removes), which is what fixes DataLoader's node and link order.
// Arithmetic is untouched: the reference's expressions are compiled as they stand, in IEEE double (-ffp-contract=off).
// One library call is not the BCL's: Math.Log is libm's log (both are faithfully rounded; they may differ in the last bit
// on rare arguments -- it only enters the MENTION weights of DataLoader.addMentionCount2).
//
// Also here: `SQLiteAdapter`, the class DataLoader calls for its data (TweetRecommender/SQLiteAdapter.cs over
// System.Data.SQLite, an un-vendored NuGet package).  Same six methods, answered from in-memory tables registered under the
// database path: each method is the one SELECT of SQLiteAdapter.cs:27-125 over rows kept in rowid order.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <limits>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <utility>
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

// System.Collections.Generic.Dictionary<K, V>; Keys / Values enumerate in insertion order (nothing is ever removed)
template <typename K, typename V>
class Dictionary {
    struct Rep {
        std::unordered_map<K, V> map;
        std::vector<K> order;
    };
    std::shared_ptr<Rep> p_ = std::make_shared<Rep>();

public:
    Dictionary() {}
    int Count() const { return (int)p_->order.size(); }
    bool ContainsKey(const K& k) const { return p_->map.find(k) != p_->map.end(); }
    void Add(const K& k, const V& v) {
        if (!p_->map.emplace(k, v).second) throw ArgumentException();
        p_->order.push_back(k);
    }
    V& operator[](const K& k) const {          // the getter (and `d[k] += x` on an existing key); never adds
        auto it = p_->map.find(k);
        if (it == p_->map.end()) throw KeyNotFoundException();
        return it->second;
    }
    std::vector<K> Keys() const { return p_->order; }
    std::vector<V> Values() const {
        std::vector<V> out;
        for (const K& k : p_->order) out.push_back(p_->map.find(k)->second);
        return out;
    }
};

// System.Collections.Generic.HashSet<T>; enumerates in insertion order (nothing is ever removed)
template <typename T>
class HashSet {
    struct Rep {
        std::unordered_set<T> set;
        std::vector<T> order;
    };
    std::shared_ptr<Rep> p_ = std::make_shared<Rep>();

public:
    HashSet() {}
    bool Add(const T& x) {
        if (!p_->set.insert(x).second) return false;
        p_->order.push_back(x);
        return true;
    }
    bool Contains(const T& x) const { return p_->set.find(x) != p_->set.end(); }
    int Count() const { return (int)p_->order.size(); }
    typename std::vector<T>::const_iterator begin() const { return p_->order.begin(); }
    typename std::vector<T>::const_iterator end() const { return p_->order.end(); }
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

// long.Parse, System.IO.Path.GetFileNameWithoutExtension, System.Math.Log
inline long long ParseLong(const std::string& s) {
    size_t used = 0;
    const long long v = std::stoll(s, &used);
    if (used != s.size()) throw std::invalid_argument("FormatException");
    return v;
}
struct Path {
    static std::string GetFileNameWithoutExtension(const std::string& path) {
        const size_t slash = path.find_last_of("/\\");
        std::string name = slash == std::string::npos ? path : path.substr(slash + 1);
        const size_t dot = name.find_last_of('.');
        return dot == std::string::npos ? name : name.substr(0, dot);
    }
};
struct Math {
    static double Log(double x) { return std::log(x); }
};

// ---- the data DataLoader asks for (TweetRecommender/SQLiteAdapter.cs) ---------------------------------------------------
struct MemDb {
    // follow(source, target), tweet(id, author), retweet(user, tweet), quote(user, tweet), favorite(user, tweet),
    // mention(source, target): rows as (first column, second column), in rowid order
    std::map<std::string, std::vector<std::pair<long long, long long>>> tables;
};
inline std::map<std::string, MemDb>& mem_dbs() {
    static std::map<std::string, MemDb> dbs;
    return dbs;
}

class SQLiteAdapter {
    const MemDb* db_;
    // SELECT <out> FROM <table> WHERE <key> = id, collected into a HashSet in row order
    HashSet<long long> select(const char* table, bool key_is_first, long long id) const {
        HashSet<long long> out;
        auto it = db_->tables.find(table);
        if (it != db_->tables.end())
            for (const auto& row : it->second)
                if ((key_is_first ? row.first : row.second) == id) out.Add(key_is_first ? row.second : row.first);
        return out;
    }

public:
    explicit SQLiteAdapter(const std::string& dbPath) {
        auto it = mem_dbs().find(dbPath);
        if (it == mem_dbs().end()) throw std::runtime_error("SQLiteException: unable to open " + dbPath);
        db_ = &it->second;
    }
    void closeDB() {}
    HashSet<long long> getFollowingUsers(long long userId) { return select("follow", true, userId); }      // :27-39  target WHERE source
    HashSet<long long> getAuthorship(long long userId) { return select("tweet", false, userId); }          // :41-53  id WHERE author
    HashSet<long long> getRetweets(long long userId) { return select("retweet", true, userId); }           // :55-67  tweet WHERE user
    HashSet<long long> getQuotedTweets(long long userId) { return select("quote", true, userId); }         // :69-81
    HashSet<long long> getFavoriteTweets(long long userId) { return select("favorite", true, userId); }    // :83-95
    int getMentionCount(long long userId1, long long userId2) {                                            // :114-125, both directions
        int count = 0;
        auto it = db_->tables.find("mention");
        if (it != db_->tables.end())
            for (const auto& row : it->second)
                count += (row.first == userId1 && row.second == userId2) + (row.first == userId2 && row.second == userId1);
        return count;
    }
};

}  // namespace bcl
