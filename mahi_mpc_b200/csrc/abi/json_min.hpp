// mahi-mpc-b200 :: a small JSON reader/writer, enough for the "<name>.json" model artefact
// (schema: src/Mahi/Mpc/ModelParameters.cpp:37-72, written by ModelGenerator.cpp:261-270).
// The reference uses nlohmann::json through mahi-util; neither is available offline.
#pragma once
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

namespace mmpc_json {

struct Value {
    enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
    bool boolean = false;
    double number = 0.0;
    std::string string;
    std::vector<Value> array;
    std::vector<std::pair<std::string, Value>> object;

    bool is_number() const { return kind == Number; }
    bool is_string() const { return kind == String; }
    bool is_array() const { return kind == Array; }
    bool is_object() const { return kind == Object; }
    bool is_bool() const { return kind == Bool; }
    const Value* get(const std::string& key) const {
        if (kind != Object) return nullptr;
        for (const auto& kv : object)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
};

class Parser {
public:
    Parser(const std::string& s) : m_s(s), m_i(0) {}
    bool parse(Value* out, std::string* err) {
        if (!value(out)) {
            if (err) *err = "JSON syntax error near offset " + std::to_string(m_i) + (m_msg.empty() ? "" : ": " + m_msg);
            return false;
        }
        ws();
        if (m_i != m_s.size()) {
            if (err) *err = "trailing characters after JSON document";
            return false;
        }
        return true;
    }

private:
    void ws() { while (m_i < m_s.size() && std::isspace(static_cast<unsigned char>(m_s[m_i]))) ++m_i; }
    bool lit(const char* w) {
        size_t n = std::char_traits<char>::length(w);
        if (m_s.compare(m_i, n, w) == 0) { m_i += n; return true; }
        return false;
    }
    bool value(Value* v) {
        ws();
        if (m_i >= m_s.size()) { m_msg = "unexpected end"; return false; }
        char c = m_s[m_i];
        if (c == '{') return object(v);
        if (c == '[') return array(v);
        if (c == '"') { v->kind = Value::String; return str(&v->string); }
        if (lit("true")) { v->kind = Value::Bool; v->boolean = true; return true; }
        if (lit("false")) { v->kind = Value::Bool; v->boolean = false; return true; }
        if (lit("null")) { v->kind = Value::Null; return true; }
        // numbers; nlohmann writes non-finite doubles as null, casadi::inf never reaches the file
        // because the sentinel is +-10e30 (ModelParameters.cpp:21-24)
        const char* b = m_s.c_str() + m_i;
        char* e = nullptr;
        double d = std::strtod(b, &e);
        if (e == b) { m_msg = "bad token"; return false; }
        m_i += static_cast<size_t>(e - b);
        v->kind = Value::Number;
        v->number = d;
        return true;
    }
    bool str(std::string* out) {
        ++m_i;
        out->clear();
        while (m_i < m_s.size() && m_s[m_i] != '"') {
            char c = m_s[m_i++];
            if (c == '\\' && m_i < m_s.size()) {
                char e = m_s[m_i++];
                switch (e) {
                    case 'n': out->push_back('\n'); break;
                    case 't': out->push_back('\t'); break;
                    case 'r': out->push_back('\r'); break;
                    case 'b': out->push_back('\b'); break;
                    case 'f': out->push_back('\f'); break;
                    case 'u': {  // keep BMP code points <= 0x7f, replace others by '?'
                        unsigned cp = 0;
                        for (int k = 0; k < 4 && m_i < m_s.size(); ++k) cp = cp * 16 + static_cast<unsigned>(std::strtol(std::string(1, m_s[m_i++]).c_str(), nullptr, 16));
                        out->push_back(cp < 0x80 ? static_cast<char>(cp) : '?');
                    } break;
                    default: out->push_back(e);
                }
            } else out->push_back(c);
        }
        if (m_i >= m_s.size()) { m_msg = "unterminated string"; return false; }
        ++m_i;
        return true;
    }
    bool array(Value* v) {
        v->kind = Value::Array;
        ++m_i;
        ws();
        if (m_i < m_s.size() && m_s[m_i] == ']') { ++m_i; return true; }
        for (;;) {
            Value e;
            if (!value(&e)) return false;
            v->array.push_back(std::move(e));
            ws();
            if (m_i < m_s.size() && m_s[m_i] == ',') { ++m_i; continue; }
            if (m_i < m_s.size() && m_s[m_i] == ']') { ++m_i; return true; }
            m_msg = "expected , or ]";
            return false;
        }
    }
    bool object(Value* v) {
        v->kind = Value::Object;
        ++m_i;
        ws();
        if (m_i < m_s.size() && m_s[m_i] == '}') { ++m_i; return true; }
        for (;;) {
            ws();
            if (m_i >= m_s.size() || m_s[m_i] != '"') { m_msg = "expected key"; return false; }
            std::string k;
            if (!str(&k)) return false;
            ws();
            if (m_i >= m_s.size() || m_s[m_i] != ':') { m_msg = "expected :"; return false; }
            ++m_i;
            Value e;
            if (!value(&e)) return false;
            v->object.emplace_back(std::move(k), std::move(e));
            ws();
            if (m_i < m_s.size() && m_s[m_i] == ',') { ++m_i; continue; }
            if (m_i < m_s.size() && m_s[m_i] == '}') { ++m_i; return true; }
            m_msg = "expected , or }";
            return false;
        }
    }
    const std::string& m_s;
    size_t m_i;
    std::string m_msg;
};

inline bool parse(const std::string& text, Value* out, std::string* err) { return Parser(text).parse(out, err); }

inline std::string quote(const std::string& s) {
    std::string o = "\"";
    for (char c : s) {
        if (c == '"' || c == '\\') { o.push_back('\\'); o.push_back(c); }
        else if (c == '\n') o += "\\n";
        else o.push_back(c);
    }
    return o + "\"";
}
// shortest representation that parses back to the same double (1e+31 rather than 9.9999999999999996e+30)
inline std::string number(double d) {
    char buf[64];
    if (d == std::floor(d) && std::fabs(d) < 1e15) {
        std::snprintf(buf, sizeof(buf), "%.1f", d);
        return buf;
    }
    for (int prec = 1; prec <= 17; ++prec) {
        std::snprintf(buf, sizeof(buf), "%.*g", prec, d);
        if (std::strtod(buf, nullptr) == d) break;
    }
    return buf;
}

}  // namespace mmpc_json
