// json.hpp -- small ordered JSON value with parser and writer (target configs in, reports out).
// juliet's target config is "a JSON file" (/root/reference/doc/JULIET.md:128-157) and its primary output is
// JSON, the HTML being "a 1:1 conversion of the JSON file" (:68-69).
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace msjson {

struct Value {
    enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
    bool b = false;
    double num = 0;
    bool is_int = false;
    std::string str;
    std::vector<Value> arr;
    std::vector<std::pair<std::string, Value>> obj;

    Value() = default;
    static Value boolean(bool v) { Value x; x.kind = Bool; x.b = v; return x; }
    static Value number(double v) { Value x; x.kind = Number; x.num = v; return x; }
    static Value integer(long long v) { Value x; x.kind = Number; x.num = static_cast<double>(v); x.is_int = true; return x; }
    static Value string(std::string v) { Value x; x.kind = String; x.str = std::move(v); return x; }
    static Value array() { Value x; x.kind = Array; return x; }
    static Value object() { Value x; x.kind = Object; return x; }

    Value& set(const std::string& k, Value v) {
        for (auto& kv : obj) if (kv.first == k) { kv.second = std::move(v); return kv.second; }
        obj.emplace_back(k, std::move(v));
        return obj.back().second;
    }
    Value& push(Value v) { arr.push_back(std::move(v)); return arr.back(); }
    const Value* get(const std::string& k) const {
        for (auto& kv : obj) if (kv.first == k) return &kv.second;
        return nullptr;
    }
    std::string get_string(const std::string& k, const std::string& dflt = "") const {
        const Value* v = get(k);
        return v && v->kind == String ? v->str : dflt;
    }
    double get_number(const std::string& k, double dflt = 0) const {
        const Value* v = get(k);
        return v && v->kind == Number ? v->num : dflt;
    }
};

class Parser {
public:
    explicit Parser(const std::string& s) : s_(s) {}
    Value parse() {
        Value v = value();
        ws();
        if (i_ != s_.size()) fail("trailing characters");
        return v;
    }
private:
    [[noreturn]] void fail(const std::string& m) { throw std::runtime_error("JSON: " + m + " at offset " + std::to_string(i_)); }
    void ws() { while (i_ < s_.size() && (s_[i_] == ' ' || s_[i_] == '\n' || s_[i_] == '\t' || s_[i_] == '\r')) ++i_; }
    Value value() {
        ws();
        if (i_ >= s_.size()) fail("unexpected end");
        const char c = s_[i_];
        if (c == '{') {
            Value o = Value::object();
            ++i_; ws();
            if (i_ < s_.size() && s_[i_] == '}') { ++i_; return o; }
            for (;;) {
                ws();
                if (i_ >= s_.size() || s_[i_] != '"') fail("expected key");
                std::string k = str();
                ws();
                if (i_ >= s_.size() || s_[i_] != ':') fail("expected ':'");
                ++i_;
                o.obj.emplace_back(k, value());
                ws();
                if (i_ < s_.size() && s_[i_] == ',') { ++i_; continue; }
                if (i_ < s_.size() && s_[i_] == '}') { ++i_; return o; }
                fail("expected ',' or '}'");
            }
        }
        if (c == '[') {
            Value a = Value::array();
            ++i_; ws();
            if (i_ < s_.size() && s_[i_] == ']') { ++i_; return a; }
            for (;;) {
                a.arr.push_back(value());
                ws();
                if (i_ < s_.size() && s_[i_] == ',') { ++i_; continue; }
                if (i_ < s_.size() && s_[i_] == ']') { ++i_; return a; }
                fail("expected ',' or ']'");
            }
        }
        if (c == '"') return Value::string(str());
        if (s_.compare(i_, 4, "true") == 0) { i_ += 4; return Value::boolean(true); }
        if (s_.compare(i_, 5, "false") == 0) { i_ += 5; return Value::boolean(false); }
        if (s_.compare(i_, 4, "null") == 0) { i_ += 4; return Value(); }
        char* end = nullptr;
        const double d = strtod(s_.c_str() + i_, &end);
        if (end == s_.c_str() + i_) fail("unexpected character");
        Value v = Value::number(d);
        const std::string tok(s_.c_str() + i_, static_cast<size_t>(end - (s_.c_str() + i_)));
        v.is_int = tok.find_first_of(".eE") == std::string::npos;
        i_ = static_cast<size_t>(end - s_.c_str());
        return v;
    }
    std::string str() {
        std::string out;
        ++i_;
        while (i_ < s_.size() && s_[i_] != '"') {
            char c = s_[i_++];
            if (c == '\\') {
                if (i_ >= s_.size()) fail("bad escape");
                const char e = s_[i_++];
                switch (e) {
                case 'n': out += '\n'; break; case 't': out += '\t'; break; case 'r': out += '\r'; break;
                case 'b': out += '\b'; break; case 'f': out += '\f'; break;
                case 'u': {
                    if (i_ + 4 > s_.size()) fail("bad \\u");
                    const unsigned cp = static_cast<unsigned>(strtoul(s_.substr(i_, 4).c_str(), nullptr, 16));
                    i_ += 4;
                    if (cp < 0x80) out += static_cast<char>(cp);
                    else if (cp < 0x800) { out += static_cast<char>(0xC0 | (cp >> 6)); out += static_cast<char>(0x80 | (cp & 0x3F)); }
                    else { out += static_cast<char>(0xE0 | (cp >> 12)); out += static_cast<char>(0x80 | ((cp >> 6) & 0x3F)); out += static_cast<char>(0x80 | (cp & 0x3F)); }
                    break;
                }
                default: out += e;
                }
            } else out += c;
        }
        if (i_ >= s_.size()) fail("unterminated string");
        ++i_;
        return out;
    }
    const std::string& s_;
    size_t i_ = 0;
};

inline void escape(const std::string& s, std::string& out) {
    out += '"';
    for (unsigned char c : s) {
        switch (c) {
        case '"': out += "\\\""; break; case '\\': out += "\\\\"; break; case '\n': out += "\\n"; break;
        case '\t': out += "\\t"; break; case '\r': out += "\\r"; break;
        default:
            if (c < 0x20) { char b[8]; snprintf(b, sizeof b, "\\u%04x", c); out += b; }
            else out += static_cast<char>(c);
        }
    }
    out += '"';
}

inline void write(const Value& v, std::string& out, int indent = 0, int step = 2) {
    const std::string pad(static_cast<size_t>(indent + step), ' '), padc(static_cast<size_t>(indent), ' ');
    switch (v.kind) {
    case Value::Null: out += "null"; break;
    case Value::Bool: out += v.b ? "true" : "false"; break;
    case Value::Number: {
        char b[40];
        if (v.is_int) snprintf(b, sizeof b, "%lld", static_cast<long long>(v.num));
        else if (std::isfinite(v.num)) snprintf(b, sizeof b, "%.17g", v.num);
        else snprintf(b, sizeof b, "null");
        out += b;
        break;
    }
    case Value::String: escape(v.str, out); break;
    case Value::Array: {
        if (v.arr.empty()) { out += "[]"; break; }
        bool scalar = true;
        for (auto& e : v.arr) if (e.kind == Value::Array || e.kind == Value::Object) scalar = false;
        if (scalar) {
            out += "[";
            for (size_t i = 0; i < v.arr.size(); ++i) { if (i) out += ", "; write(v.arr[i], out, indent, step); }
            out += "]";
        } else {
            out += "[\n";
            for (size_t i = 0; i < v.arr.size(); ++i) { out += pad; write(v.arr[i], out, indent + step, step); out += i + 1 < v.arr.size() ? ",\n" : "\n"; }
            out += padc + "]";
        }
        break;
    }
    case Value::Object: {
        if (v.obj.empty()) { out += "{}"; break; }
        out += "{\n";
        for (size_t i = 0; i < v.obj.size(); ++i) {
            out += pad; escape(v.obj[i].first, out); out += ": ";
            write(v.obj[i].second, out, indent + step, step);
            out += i + 1 < v.obj.size() ? ",\n" : "\n";
        }
        out += padc + "}";
        break;
    }
    }
}

inline std::string dump(const Value& v) { std::string s; write(v, s); s += "\n"; return s; }

}  // namespace msjson
