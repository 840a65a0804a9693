// mahi-mpc-b200 :: the slice of mahi::util the MPC library and its drivers use (the real mahi-util is
// FetchContent'ed unpinned by the reference, CMakeLists.txt:42, and is not available offline).
// Surface: Time + seconds/milliseconds/microseconds, Clock, Timer, sleep, PI, mean, print,
// Options (cxxopts-style flags used by the example drivers).
#pragma once
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <iostream>
#include <map>
#include <numeric>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

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

/// Minimal command-line flags in the style of mahi::util::Options (cxxopts): "s,long" names, count(), values.
class Options {
public:
    class Result {
    public:
        size_t count(const std::string& name) const { auto it = m_vals.find(name); return it == m_vals.end() ? 0 : it->second.size(); }
        const std::vector<std::string>& values(const std::string& name) const {
            static const std::vector<std::string> none;
            auto it = m_vals.find(name);
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
        std::map<std::string, std::vector<std::string>> m_vals;
    };
    class Adder {
    public:
        explicit Adder(Options& o) : m_o(o) {}
        Adder& operator()(const std::string& names, const std::string& desc, bool takes_value = false) {
            m_o.add(names, desc, takes_value);
            return *this;
        }
    private:
        Options& m_o;
    };
    Options(std::string program, std::string help) : m_program(std::move(program)), m_help(std::move(help)) {}
    Adder add_options() { return Adder(*this); }
    Result parse(int argc, char* argv[]) const {
        Result r;
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
