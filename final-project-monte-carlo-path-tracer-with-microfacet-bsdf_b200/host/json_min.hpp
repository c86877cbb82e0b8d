// host/json_min.hpp — a small recursive-descent JSON reader for conf.json.
//
// The reference reads conf.json through nlohmann::json (src/main.cpp:147-289).  Only the
// queries main.cpp makes are needed here: object/array access, is_number / is_number_float /
// is_boolean / is_string / is_null, contains(), and conversion to int/float/bool/string with a
// type error (exception) when the value has another type — main.cpp's single try/catch turns
// such an error into "stop reading the rest of the file" (src/main.cpp:291-294), which the
// scene assembler reproduces.
#pragma once
#include <cmath>
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace b2pt_host {

struct JsonError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

class Json {
  public:
    enum Kind { Null, Bool, Int, Float, String, Array, Object };
    Kind kind = Null;
    bool b = false;
    long long i = 0;
    double d = 0.0;
    std::string s;
    std::vector<Json> arr;
    std::vector<std::pair<std::string, Json>> obj;  // insertion order kept; lookups are linear (tiny files)

    bool is_null() const { return kind == Null; }
    bool is_boolean() const { return kind == Bool; }
    bool is_number() const { return kind == Int || kind == Float; }
    bool is_number_float() const { return kind == Float; }
    bool is_string() const { return kind == String; }
    bool is_array() const { return kind == Array; }
    bool is_object() const { return kind == Object; }
    size_t size() const { return kind == Array ? arr.size() : kind == Object ? obj.size() : (kind == Null ? 0 : 1); }

    bool contains(const std::string &key) const {
        if (kind != Object) return false;
        for (auto &kv : obj)
            if (kv.first == key) return true;
        return false;
    }
    // operator[] on a missing key yields null (nlohmann on a non-const copy inserts null).
    const Json &operator[](const std::string &key) const {
        static const Json null_value;
        if (kind == Null) return null_value;
        if (kind != Object) throw JsonError("cannot use operator[] with a string argument with a non-object");
        for (auto &kv : obj)
            if (kv.first == key) return kv.second;
        return null_value;
    }
    const Json &operator[](size_t idx) const {
        if (kind != Array) throw JsonError("cannot use operator[] with a numeric argument with a non-array");
        if (idx >= arr.size()) throw JsonError("array index out of range");
        return arr[idx];
    }
    double as_double() const {
        if (kind == Int) return (double)i;
        if (kind == Float) return d;
        if (kind == Bool) return b ? 1.0 : 0.0;  // nlohmann converts booleans to arithmetic types
        throw JsonError("type must be number, but is " + kind_name());
    }
    float as_float() const { return (float)as_double(); }
    int as_int() const {
        if (kind == Int) return (int)i;
        if (kind == Float) return (int)d;  // static_cast truncation, like nlohmann's get<int>()
        if (kind == Bool) return b ? 1 : 0;
        throw JsonError("type must be number, but is " + kind_name());
    }
    bool as_bool() const {
        if (kind != Bool) throw JsonError("type must be boolean, but is " + kind_name());
        return b;
    }
    const std::string &as_string() const {
        if (kind != String) throw JsonError("type must be string, but is " + kind_name());
        return s;
    }
    std::string kind_name() const {
        switch (kind) {
        case Null: return "null";
        case Bool: return "boolean";
        case Int: case Float: return "number";
        case String: return "string";
        case Array: return "array";
        default: return "object";
        }
    }

    static Json parse(const std::string &text) {
        Parser p{text, 0};
        p.ws();
        Json v = p.value();
        p.ws();
        if (p.pos != text.size()) throw JsonError("parse error: trailing characters at offset " + std::to_string(p.pos));
        return v;
    }

  private:
    struct Parser {
        const std::string &t;
        size_t pos;
        void ws() {
            while (pos < t.size() && (t[pos] == ' ' || t[pos] == '\t' || t[pos] == '\n' || t[pos] == '\r')) ++pos;
        }
        [[noreturn]] void fail(const char *what) {
            throw JsonError(std::string("parse error at offset ") + std::to_string(pos) + ": " + what);
        }
        bool lit(const char *w) {
            size_t n = std::char_traits<char>::length(w);
            if (t.compare(pos, n, w) == 0) { pos += n; return true; }
            return false;
        }
        Json value() {
            if (pos >= t.size()) fail("unexpected end of input");
            char c = t[pos];
            Json v;
            if (c == '{') {
                v.kind = Object; ++pos; ws();
                if (pos < t.size() && t[pos] == '}') { ++pos; return v; }
                for (;;) {
                    ws();
                    if (pos >= t.size() || t[pos] != '"') fail("expected string key");
                    std::string key = str();
                    ws();
                    if (pos >= t.size() || t[pos] != ':') fail("expected ':'");
                    ++pos; ws();
                    Json val = value();
                    bool replaced = false;
                    for (auto &kv : v.obj)
                        if (kv.first == key) { kv.second = val; replaced = true; }
                    if (!replaced) v.obj.emplace_back(key, std::move(val));
                    ws();
                    if (pos < t.size() && t[pos] == ',') { ++pos; continue; }
                    if (pos < t.size() && t[pos] == '}') { ++pos; break; }
                    fail("expected ',' or '}'");
                }
                return v;
            }
            if (c == '[') {
                v.kind = Array; ++pos; ws();
                if (pos < t.size() && t[pos] == ']') { ++pos; return v; }
                for (;;) {
                    ws();
                    v.arr.push_back(value());
                    ws();
                    if (pos < t.size() && t[pos] == ',') { ++pos; continue; }
                    if (pos < t.size() && t[pos] == ']') { ++pos; break; }
                    fail("expected ',' or ']'");
                }
                return v;
            }
            if (c == '"') { v.kind = String; v.s = str(); return v; }
            if (lit("true")) { v.kind = Bool; v.b = true; return v; }
            if (lit("false")) { v.kind = Bool; v.b = false; return v; }
            if (lit("null")) return v;
            if (c == '-' || (c >= '0' && c <= '9')) return number();
            fail("unexpected character");
        }
        std::string str() {
            std::string out;
            ++pos;  // opening quote
            while (pos < t.size() && t[pos] != '"') {
                char c = t[pos++];
                if (c == '\\') {
                    if (pos >= t.size()) fail("bad escape");
                    char e = t[pos++];
                    switch (e) {
                    case 'n': out += '\n'; break;
                    case 't': out += '\t'; break;
                    case 'r': out += '\r'; break;
                    case 'b': out += '\b'; break;
                    case 'f': out += '\f'; break;
                    case 'u': {
                        if (pos + 4 > t.size()) fail("bad \\u escape");
                        unsigned cp = (unsigned)std::strtoul(t.substr(pos, 4).c_str(), nullptr, 16);
                        pos += 4;
                        if (cp < 0x80) out += (char)cp;
                        else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
                        else { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
                        break;
                    }
                    default: out += e;  // \" \\ \/
                    }
                } else out += c;
            }
            if (pos >= t.size()) fail("unterminated string");
            ++pos;
            return out;
        }
        Json number() {
            size_t start = pos;
            bool is_float = false;
            if (t[pos] == '-') ++pos;
            while (pos < t.size() && t[pos] >= '0' && t[pos] <= '9') ++pos;
            if (pos < t.size() && t[pos] == '.') { is_float = true; ++pos; while (pos < t.size() && t[pos] >= '0' && t[pos] <= '9') ++pos; }
            if (pos < t.size() && (t[pos] == 'e' || t[pos] == 'E')) {
                is_float = true; ++pos;
                if (pos < t.size() && (t[pos] == '+' || t[pos] == '-')) ++pos;
                while (pos < t.size() && t[pos] >= '0' && t[pos] <= '9') ++pos;
            }
            std::string tok = t.substr(start, pos - start);
            Json v;
            if (is_float) { v.kind = Float; v.d = std::strtod(tok.c_str(), nullptr); }
            else { v.kind = Int; v.i = std::strtoll(tok.c_str(), nullptr, 10); v.d = (double)v.i; }
            return v;
        }
    };
};

}  // namespace b2pt_host
