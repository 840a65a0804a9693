// mahi-mpc-b200 :: mahi::util::json -- the slice of nlohmann::json (re-exported by mahi-util as mahi::util::json) that
// mahi::mpc and its drivers use: the "<name>.json" model artefact written by ModelGenerator::save_param_file
// (src/Mahi/Mpc/ModelGenerator.cpp:261-270: `j["model"] = params; file << j;`), read back by ModelControl::load_model
// (src/Mahi/Mpc/ModelControl.cpp:22-26: `file >> j; j["model"].get<ModelParameters>()`), and the to_json / from_json
// pair of include/Mahi/Mpc/ModelParameters.hpp:43-45 (brace-initialised object, at(), get_to(), implicit conversion of a
// number).  nlohmann/json is not available offline; this is a small self-contained value type with the same spelling.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <initializer_list>
#include <istream>
#include <iterator>
#include <map>
#include <ostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

namespace mahi {
namespace util {

class json {
public:
    enum class kind { null, boolean, integer, real, string, array, object };
    using object_t = std::map<std::string, json>;  // sorted keys, like nlohmann's default
    using array_t = std::vector<json>;

    json() = default;
    json(std::nullptr_t) {}
    json(bool v) : m_kind(kind::boolean), m_bool(v) {}
    template <class T, typename std::enable_if<std::is_integral<T>::value && !std::is_same<T, bool>::value, int>::type = 0>
    json(T v) : m_kind(kind::integer), m_int(static_cast<std::int64_t>(v)) {}
    template <class T, typename std::enable_if<std::is_floating_point<T>::value, int>::type = 0>
    json(T v) : m_kind(kind::real), m_real(static_cast<double>(v)) {}
    json(const char* s) : m_kind(kind::string), m_str(s) {}
    json(const std::string& s) : m_kind(kind::string), m_str(s) {}
    template <class T>
    json(const std::vector<T>& v) : m_kind(kind::array) {
        for (const T& e : v) m_arr.emplace_back(e);
    }
    /// json{{"key", value}, ...} -> object (every element a [string, value] pair); anything else -> array
    json(std::initializer_list<json> init) {
        bool pairs = init.size() > 0;
        for (const json& e : init) pairs = pairs && e.m_kind == kind::array && e.m_arr.size() == 2 && e.m_arr[0].m_kind == kind::string;
        if (pairs) {
            m_kind = kind::object;
            for (const json& e : init) m_obj[e.m_arr[0].m_str] = e.m_arr[1];
        } else {
            m_kind = kind::array;
            m_arr.assign(init.begin(), init.end());
        }
    }
    /// any type with an ADL-visible `void to_json(json&, const T&)` (e.g. mahi::mpc::ModelParameters)
    template <class T, typename std::enable_if<!std::is_arithmetic<T>::value && !std::is_convertible<T, std::string>::value &&
                                                   !std::is_same<typename std::decay<T>::type, json>::value,
                                               decltype(to_json(std::declval<json&>(), std::declval<const T&>()), 0)>::type = 0>
    json(const T& v) {
        to_json(*this, v);
    }

    kind type() const { return m_kind; }
    bool is_null() const { return m_kind == kind::null; }
    bool is_object() const { return m_kind == kind::object; }
    bool is_array() const { return m_kind == kind::array; }
    bool is_string() const { return m_kind == kind::string; }
    bool is_boolean() const { return m_kind == kind::boolean; }
    bool is_number() const { return m_kind == kind::integer || m_kind == kind::real; }
    size_t size() const { return m_kind == kind::array ? m_arr.size() : (m_kind == kind::object ? m_obj.size() : (m_kind == kind::null ? 0 : 1)); }
    bool contains(const std::string& key) const { return m_kind == kind::object && m_obj.count(key) != 0; }

    json& operator[](const std::string& key) {
        if (m_kind == kind::null) m_kind = kind::object;
        if (m_kind != kind::object) throw std::domain_error("json: operator[] with a string key on a non-object");
        return m_obj[key];
    }
    const json& operator[](const std::string& key) const { return at(key); }
    json& operator[](const char* key) { return (*this)[std::string(key)]; }
    const json& operator[](const char* key) const { return at(std::string(key)); }
    const json& at(const std::string& key) const {
        if (m_kind != kind::object) throw std::domain_error("json: at() with a string key on a non-object");
        auto it = m_obj.find(key);
        if (it == m_obj.end()) throw std::out_of_range("json: key '" + key + "' not found");
        return it->second;
    }
    json& at(const std::string& key) { return const_cast<json&>(static_cast<const json&>(*this).at(key)); }
    const json& at(size_t i) const {
        if (m_kind != kind::array || i >= m_arr.size()) throw std::out_of_range("json: array index out of range");
        return m_arr[i];
    }
    void push_back(const json& v) {
        if (m_kind == kind::null) m_kind = kind::array;
        if (m_kind != kind::array) throw std::domain_error("json: push_back on a non-array");
        m_arr.push_back(v);
    }

    // ---- conversions ----
    template <class T>
    T get() const {
        T v{};
        get_to(v);
        return v;
    }
    void get_to(bool& v) const {
        if (m_kind != kind::boolean) throw std::domain_error("json: type must be boolean");
        v = m_bool;
    }
    template <class T, typename std::enable_if<std::is_arithmetic<T>::value && !std::is_same<T, bool>::value, int>::type = 0>
    void get_to(T& v) const {
        if (m_kind == kind::integer) v = static_cast<T>(m_int);
        else if (m_kind == kind::real) v = static_cast<T>(m_real);
        else if (m_kind == kind::boolean) v = static_cast<T>(m_bool);
        else throw std::domain_error("json: type must be number");
    }
    void get_to(std::string& v) const {
        if (m_kind != kind::string) throw std::domain_error("json: type must be string");
        v = m_str;
    }
    template <class T>
    void get_to(std::vector<T>& v) const {
        if (m_kind != kind::array) throw std::domain_error("json: type must be array");
        v.clear();
        for (const json& e : m_arr) v.push_back(e.get<T>());
    }
    template <class T, typename std::enable_if<!std::is_arithmetic<T>::value && !std::is_same<T, std::string>::value,
                                               decltype(from_json(std::declval<const json&>(), std::declval<T&>()), 0)>::type = 0>
    void get_to(T& v) const {
        from_json(*this, v);
    }
    /// implicit conversion of a number (the reference writes `microseconds(j.at("timespan"))`, ModelParameters.cpp:54-55)
    template <class T, typename std::enable_if<std::is_arithmetic<T>::value, int>::type = 0>
    operator T() const {  // NOLINT
        return get<T>();
    }
    operator std::string() const { return get<std::string>(); }  // NOLINT

    // ---- text ----
    std::string dump() const {
        std::string out;
        write(out);
        return out;
    }
    static json parse(const std::string& text) {
        size_t i = 0;
        json v = parse_value(text, i);
        skip_ws(text, i);
        if (i != text.size()) throw std::invalid_argument("json: trailing characters at offset " + std::to_string(i));
        return v;
    }
    friend std::ostream& operator<<(std::ostream& os, const json& j) { return os << j.dump(); }
    friend std::istream& operator>>(std::istream& is, json& j) {
        std::string text((std::istreambuf_iterator<char>(is)), std::istreambuf_iterator<char>());
        j = parse(text);
        return is;
    }

private:
    kind m_kind = kind::null;
    bool m_bool = false;
    std::int64_t m_int = 0;
    double m_real = 0.0;
    std::string m_str;
    array_t m_arr;
    object_t m_obj;

    static void write_string(std::string& out, const std::string& s) {
        out += '"';
        for (unsigned char c : s) {
            switch (c) {
                case '"': out += "\\\""; break;
                case '\\': out += "\\\\"; break;
                case '\n': out += "\\n"; break;
                case '\r': out += "\\r"; break;
                case '\t': out += "\\t"; break;
                default:
                    if (c < 0x20) { char b[8]; std::snprintf(b, sizeof b, "\\u%04x", c); out += b; }
                    else out += static_cast<char>(c);
            }
        }
        out += '"';
    }
    void write(std::string& out) const {
        switch (m_kind) {
            case kind::null: out += "null"; break;
            case kind::boolean: out += m_bool ? "true" : "false"; break;
            case kind::integer: out += std::to_string(m_int); break;
            case kind::real: {
                if (!std::isfinite(m_real)) { out += "null"; break; }  // nlohmann writes non-finite numbers as null
                char b[40];
                std::snprintf(b, sizeof b, "%.17g", m_real);  // round-trips; keeps a ".0" so the value reads back as a real
                std::string s(b);
                if (s.find_first_of(".eEn") == std::string::npos) s += ".0";
                out += s;
                break;
            }
            case kind::string: write_string(out, m_str); break;
            case kind::array: {
                out += '[';
                for (size_t i = 0; i < m_arr.size(); ++i) { if (i) out += ','; m_arr[i].write(out); }
                out += ']';
                break;
            }
            case kind::object: {
                out += '{';
                bool first = true;
                for (const auto& kv : m_obj) {
                    if (!first) out += ',';
                    first = false;
                    write_string(out, kv.first);
                    out += ':';
                    kv.second.write(out);
                }
                out += '}';
                break;
            }
        }
    }
    static void skip_ws(const std::string& s, size_t& i) {
        while (i < s.size() && (s[i] == ' ' || s[i] == '\t' || s[i] == '\n' || s[i] == '\r')) ++i;
    }
    [[noreturn]] static void fail(const std::string& what, size_t i) { throw std::invalid_argument("json: " + what + " at offset " + std::to_string(i)); }
    static std::string parse_string(const std::string& s, size_t& i) {
        std::string out;
        ++i;  // opening quote
        while (true) {
            if (i >= s.size()) fail("unterminated string", i);
            char c = s[i++];
            if (c == '"') return out;
            if (c != '\\') { out += c; continue; }
            if (i >= s.size()) fail("bad escape", i);
            char e = s[i++];
            switch (e) {
                case '"': out += '"'; break;
                case '\\': out += '\\'; break;
                case '/': out += '/'; break;
                case 'b': out += '\b'; break;
                case 'f': out += '\f'; break;
                case 'n': out += '\n'; break;
                case 'r': out += '\r'; break;
                case 't': out += '\t'; break;
                case 'u': {
                    if (i + 4 > s.size()) fail("bad \\u escape", i);
                    unsigned cp = static_cast<unsigned>(std::strtoul(s.substr(i, 4).c_str(), nullptr, 16));
                    i += 4;
                    if (cp < 0x80) out += static_cast<char>(cp);
                    else if (cp < 0x800) { out += static_cast<char>(0xC0 | (cp >> 6)); out += static_cast<char>(0x80 | (cp & 0x3F)); }
                    else { out += static_cast<char>(0xE0 | (cp >> 12)); out += static_cast<char>(0x80 | ((cp >> 6) & 0x3F)); out += static_cast<char>(0x80 | (cp & 0x3F)); }
                    break;
                }
                default: fail("bad escape", i);
            }
        }
    }
    static json parse_value(const std::string& s, size_t& i) {
        skip_ws(s, i);
        if (i >= s.size()) fail("unexpected end", i);
        const char c = s[i];
        if (c == '{') {
            json o;
            o.m_kind = kind::object;
            ++i;
            skip_ws(s, i);
            if (i < s.size() && s[i] == '}') { ++i; return o; }
            while (true) {
                skip_ws(s, i);
                if (i >= s.size() || s[i] != '"') fail("expected a key", i);
                std::string key = parse_string(s, i);
                skip_ws(s, i);
                if (i >= s.size() || s[i] != ':') fail("expected ':'", i);
                ++i;
                o.m_obj[key] = parse_value(s, i);
                skip_ws(s, i);
                if (i < s.size() && s[i] == ',') { ++i; continue; }
                if (i < s.size() && s[i] == '}') { ++i; return o; }
                fail("expected ',' or '}'", i);
            }
        }
        if (c == '[') {
            json a;
            a.m_kind = kind::array;
            ++i;
            skip_ws(s, i);
            if (i < s.size() && s[i] == ']') { ++i; return a; }
            while (true) {
                a.m_arr.push_back(parse_value(s, i));
                skip_ws(s, i);
                if (i < s.size() && s[i] == ',') { ++i; continue; }
                if (i < s.size() && s[i] == ']') { ++i; return a; }
                fail("expected ',' or ']'", i);
            }
        }
        if (c == '"') return json(parse_string(s, i));
        if (s.compare(i, 4, "true") == 0) { i += 4; return json(true); }
        if (s.compare(i, 5, "false") == 0) { i += 5; return json(false); }
        if (s.compare(i, 4, "null") == 0) { i += 4; return json(); }
        const char* b = s.c_str() + i;
        char* e = nullptr;
        const double d = std::strtod(b, &e);
        if (e == b) fail("bad token", i);
        const std::string tok(b, static_cast<size_t>(e - b));
        i += tok.size();
        if (tok.find_first_of(".eE") == std::string::npos && std::fabs(d) < 9.2e18) return json(static_cast<std::int64_t>(std::strtoll(tok.c_str(), nullptr, 10)));
        return json(d);
    }
};

}  // namespace util
}  // namespace mahi

namespace nlohmann {
using json = mahi::util::json;  // ModelGenerator.cpp:262 spells it nlohmann::json
}
