// mahi-mpc-b200 :: the slice of mahi::util the MPC library and its drivers use (the real mahi-util is
// FetchContent'ed unpinned by the reference, CMakeLists.txt:42, and is not available offline).
// Surface: Time + seconds/milliseconds/microseconds, Clock, Timer, Timestamp, sleep, PI, mean, print,
// json (Mahi/Util/Json.hpp), Options (cxxopts-style flags incl. value<T>() / result["x"].as<T>() as used by
// examples/thread_model_control_example.cpp:13-28), stream output of std::vector (examples/model_control_example.cpp:104).
#pragma once
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <ctime>
#include <fstream>
#include <iostream>
#include <memory>
#include <map>
#include <numeric>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include <Mahi/Util/Json.hpp>

/// `file << std::vector<T>` as the reference's drivers write it (mahi-util provides this overload at global scope);
/// prints "[a, b, c]" -- valid MATLAB, which is what the drivers emit.
template <class T>
std::ostream& operator<<(std::ostream& os, const std::vector<T>& v) {
    os << "[";
    for (size_t i = 0; i < v.size(); ++i) os << (i ? ", " : "") << v[i];
    return os << "]";
}

namespace mahi {
namespace util {

constexpr double PI = 3.14159265358979323846;

/// Time value with microsecond resolution (mahi::util::Time look-alike).
class Time {
public:
    constexpr Time() : m_us(0) {}
    constexpr double as_seconds() const { return static_cast<double>(m_us) * 1e-6; }
    constexpr std::int32_t as_milliseconds() const { return static_cast<std::int32_t>(m_us / 1000); }
    constexpr std::int64_t as_microseconds() const { return m_us; }
    static constexpr Time from_us(std::int64_t us) { Time t; t.m_us = us; return t; }
    friend constexpr bool operator<(Time a, Time b) { return a.m_us < b.m_us; }
    friend constexpr bool operator>(Time a, Time b) { return a.m_us > b.m_us; }
    friend constexpr bool operator<=(Time a, Time b) { return a.m_us <= b.m_us; }
    friend constexpr bool operator>=(Time a, Time b) { return a.m_us >= b.m_us; }
    friend constexpr bool operator==(Time a, Time b) { return a.m_us == b.m_us; }
    friend constexpr Time operator+(Time a, Time b) { return from_us(a.m_us + b.m_us); }
    friend constexpr Time operator-(Time a, Time b) { return from_us(a.m_us - b.m_us); }
private:
    std::int64_t m_us;
};
inline Time seconds(double s) { return Time::from_us(static_cast<std::int64_t>(std::llround(s * 1e6))); }
inline Time milliseconds(std::int32_t ms) { return Time::from_us(static_cast<std::int64_t>(ms) * 1000); }
inline Time microseconds(std::int64_t us) { return Time::from_us(us); }

class Clock {
public:
    Clock() : m_t0(std::chrono::steady_clock::now()) {}
    Time get_elapsed_time() const {
        return microseconds(std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - m_t0).count());
    }
    Time restart() { Time t = get_elapsed_time(); m_t0 = std::chrono::steady_clock::now(); return t; }
private:
    std::chrono::steady_clock::time_point m_t0;
};

/// Wall-clock time stamp (mahi::util::Timestamp look-alike; the drivers print ts.hh_mm_ss_mmm() into their result files).
class Timestamp {
public:
    Timestamp() {
        const auto now = std::chrono::system_clock::now();
        const std::time_t t = std::chrono::system_clock::to_time_t(now);
        std::tm tmv{};
        localtime_r(&t, &tmv);
        year = tmv.tm_year + 1900; month = tmv.tm_mon + 1; yday = tmv.tm_yday; mday = tmv.tm_mday; wday = tmv.tm_wday;
        hour = tmv.tm_hour; min = tmv.tm_min; sec = tmv.tm_sec;
        millis = static_cast<int>(std::chrono::duration_cast<std::chrono::milliseconds>(now.time_since_epoch()).count() % 1000);
    }
    std::string yyyy_mm_dd() const { return fmt("%04d-%02d-%02d", year, month, mday); }
    std::string hh_mm_ss() const { return fmt("%02d:%02d:%02d", hour, min, sec); }
    std::string hh_mm_ss_mmm() const { return hh_mm_ss() + fmt(".%03d", millis, 0, 0); }
    std::string yyyy_mm_dd_hh_mm_ss() const { return yyyy_mm_dd() + "_" + hh_mm_ss(); }
    int year, month, yday, mday, wday, hour, min, sec, millis;
private:
    static std::string fmt(const char* f, int a, int b, int c) {
        char buf[48];
        std::snprintf(buf, sizeof buf, f, a, b, c);
        return buf;
    }
};

inline void sleep(Time t) { std::this_thread::sleep_for(std::chrono::microseconds(t.as_microseconds())); }

/// Fixed-rate loop timer: wait() blocks until the next period boundary and returns the elapsed time.
class Timer {
public:
    explicit Timer(Time period) : m_period(period), m_ticks(0) {}
    Time wait() {
        ++m_ticks;
        const Time target = microseconds(m_period.as_microseconds() * m_ticks);
        const Time now = m_clock.get_elapsed_time();
        if (now < target) sleep(target - now);
        return m_clock.get_elapsed_time();
    }
    Time get_elapsed_time() const { return m_clock.get_elapsed_time(); }
private:
    Time m_period;
    std::int64_t m_ticks;
    Clock m_clock;
};

inline double mean(const std::vector<double>& v) { return v.empty() ? 0.0 : std::accumulate(v.begin(), v.end(), 0.0) / static_cast<double>(v.size()); }

namespace detail {
inline void format_one(std::ostringstream& os, const std::string& spec, double v) {
    if (spec.size() >= 3 && spec[0] == ':' && spec[1] == '.') {
        char buf[64];
        std::snprintf(buf, sizeof(buf), ("%" + spec.substr(1)).c_str(), v);
        os << buf;
    } else os << v;
}
template <class T>
typename std::enable_if<std::is_arithmetic<T>::value>::type format_arg(std::ostringstream& os, const std::string& spec, const T& v) {
    if (!spec.empty()) format_one(os, spec, static_cast<double>(v)); else os << v;
}
template <class T>
typename std::enable_if<!std::is_arithmetic<T>::value>::type format_arg(std::ostringstream& os, const std::string&, const T& v) { os << v; }
inline void format_rest(std::ostringstream& os, const std::string& f, size_t pos) { os << f.substr(pos); }
template <class T, class... R>
void format_rest(std::ostringstream& os, const std::string& f, size_t pos, const T& v, const R&... rest) {
    const size_t a = f.find('{', pos);
    if (a == std::string::npos) { os << f.substr(pos); return; }
    const size_t b = f.find('}', a);
    os << f.substr(pos, a - pos);
    format_arg(os, f.substr(a + 1, b - a - 1), v);
    format_rest(os, f, b + 1, rest...);
}
}  // namespace detail

/// print("text {} {:.2f}", a, b) -- the fmt-style helper the reference's drivers call.
template <class... A>
void print(const std::string& f, const A&... args) {
    std::ostringstream os;
    detail::format_rest(os, f, 0, args...);
    std::cout << os.str() << std::endl;
}

/// value<T>() marks an option that takes an argument of type T (cxxopts spelling)
struct ValueSpec {};
template <class T>
std::shared_ptr<ValueSpec> value() { return std::make_shared<ValueSpec>(); }

/// Minimal command-line flags in the style of mahi::util::Options (cxxopts): "s,long" names, count(), values,
/// result["name"].as<T>() for T = arithmetic, std::string, std::vector<arithmetic> (comma separated and/or repeated).
class Options {
public:
    class OptionValue {
    public:
        explicit OptionValue(std::vector<std::string> raw) : m_raw(std::move(raw)) {}
        size_t count() const { return m_raw.size(); }
        template <class T>
        T as() const { T out{}; convert(out); return out; }
    private:
        static void parse_one(const std::string& s, std::string& out) { out = s; }
        static void parse_one(const std::string& s, bool& out) { out = !(s == "0" || s == "false" || s == "False"); }
        template <class T>
        static typename std::enable_if<std::is_arithmetic<T>::value>::type parse_one(const std::string& s, T& out) {
            std::stringstream ss(s);
            double d = 0;
            if (!(ss >> d)) throw std::invalid_argument("option value '" + s + "' is not a number");
            out = static_cast<T>(d);
        }
        template <class T>
        void convert(T& out) const {
            if (m_raw.empty()) throw std::domain_error("option has no value");
            parse_one(m_raw.back(), out);
        }
        template <class T>
        void convert(std::vector<T>& out) const {
            for (const auto& s : m_raw) {
                std::stringstream ss(s);
                std::string tok;
                while (std::getline(ss, tok, ',')) if (!tok.empty()) { T v{}; parse_one(tok, v); out.push_back(v); }
            }
        }
        std::vector<std::string> m_raw;
    };
    class Result {
    public:
        OptionValue operator[](const std::string& name) const { return OptionValue(values(name)); }
        size_t count(const std::string& name) const { auto it = m_vals.find(key(name)); return it == m_vals.end() ? 0 : it->second.size(); }
        const std::vector<std::string>& values(const std::string& name) const {
            static const std::vector<std::string> none;
            auto it = m_vals.find(key(name));
            return it == m_vals.end() ? none : it->second;
        }
        std::vector<double> doubles(const std::string& name) const {
            std::vector<double> out;
            for (const auto& s : values(name)) {
                std::stringstream ss(s);
                std::string tok;
                while (std::getline(ss, tok, ',')) if (!tok.empty()) out.push_back(std::stod(tok));
            }
            return out;
        }
    private:
        friend class Options;
        const std::string& key(const std::string& name) const { auto it = m_short.find(name); return it == m_short.end() ? name : it->second; }
        std::map<std::string, std::vector<std::string>> m_vals;
        std::map<std::string, std::string> m_short;  // short name -> long name: count("m") == count("ma27")
    };
    class Adder {
    public:
        explicit Adder(Options& o) : m_o(o) {}
        Adder& operator()(const std::string& names, const std::string& desc, bool takes_value = false) {
            m_o.add(names, desc, takes_value);
            return *this;
        }
        Adder& operator()(const std::string& names, const std::string& desc, const std::shared_ptr<ValueSpec>&) {
            m_o.add(names, desc, true);
            return *this;
        }
    private:
        Options& m_o;
    };
    Options(std::string program, std::string help) : m_program(std::move(program)), m_help(std::move(help)) {}
    Adder add_options() { return Adder(*this); }
    Result parse(int argc, char* argv[]) const {
        Result r;
        r.m_short = m_short;
        for (int i = 1; i < argc; ++i) {
            std::string a = argv[i];
            std::string key;
            if (a.rfind("--", 0) == 0) key = a.substr(2);
            else if (a.rfind("-", 0) == 0 && a.size() >= 2) { auto it = m_short.find(a.substr(1, 1)); key = it == m_short.end() ? a.substr(1) : it->second; }
            else continue;
            std::string val;
            const size_t eq = key.find('=');
            if (eq != std::string::npos) { val = key.substr(eq + 1); key = key.substr(0, eq); }
            auto tv = m_takes.find(key);
            if (tv == m_takes.end()) throw std::invalid_argument("unknown option --" + key);
            if (tv->second && val.empty() && i + 1 < argc) val = argv[++i];
            r.m_vals[key].push_back(val);
        }
        return r;
    }
    std::string help() const {
        std::ostringstream os;
        os << m_help << "\nusage: " << m_program << " [options]\n";
        for (const auto& d : m_desc) os << "  " << d << "\n";
        return os.str();
    }
private:
    void add(const std::string& names, const std::string& desc, bool takes_value) {
        std::string s, l = names;
        const size_t c = names.find(',');
        if (c != std::string::npos) { s = names.substr(0, c); l = names.substr(c + 1); }
        auto trim = [](std::string& t) {  // the reference writes "q, q_vec" (thread_model_control_example.cpp:16)
            while (!t.empty() && t.front() == ' ') t.erase(t.begin());
            while (!t.empty() && t.back() == ' ') t.pop_back();
        };
        trim(s); trim(l);
        if (!s.empty()) m_short[s] = l;
        m_takes[l] = takes_value;
        m_desc.push_back((s.empty() ? "" : "-" + s + ", ") + "--" + l + "  " + desc);
    }
    std::string m_program, m_help;
    std::map<std::string, std::string> m_short;
    std::map<std::string, bool> m_takes;
    std::vector<std::string> m_desc;
};

}  // namespace util
}  // namespace mahi
