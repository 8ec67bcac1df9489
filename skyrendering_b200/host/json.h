// Minimal JSON DOM for the scene files.  Accepts what the reference's loader accepts:
// rapidjson with kParseCommentsFlag | kParseTrailingCommasFlag (src/SkyRendering/AppWindow.cpp:33-36).
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace skyhost {

struct Json {
    enum Type { Null, Bool, Number, String, Array, Object } type = Null;
    bool b = false;
    bool is_int = false;
    double num = 0.0;
    std::string str;
    std::vector<Json> arr;
    std::vector<std::pair<std::string, Json>> obj;  // insertion order kept

    const Json* find(const std::string& key) const {
        if (type != Object) return nullptr;
        for (const auto& kv : obj) if (kv.first == key) return &kv.second;
        return nullptr;
    }
    Json& set(const std::string& key) {
        type = Object;
        for (auto& kv : obj) if (kv.first == key) return kv.second;
        obj.emplace_back(key, Json());
        return obj.back().second;
    }
    static Json number(double v) { Json j; j.type = Number; j.num = v; return j; }
    static Json integer(long long v) { Json j; j.type = Number; j.num = double(v); j.is_int = true; return j; }
    static Json string(const std::string& s) { Json j; j.type = String; j.str = s; return j; }
    static Json boolean(bool v) { Json j; j.type = Bool; j.b = v; return j; }
};

class JsonParser {
public:
    explicit JsonParser(const std::string& text) : s_(text) {}
    Json parse() {
        Json v = value();
        skip();
        if (p_ != s_.size()) error("trailing characters");
        return v;
    }

private:
    const std::string& s_;
    size_t p_ = 0;

    [[noreturn]] void error(const char* what) const {
        throw std::runtime_error("JSON parse error (offset " + std::to_string(p_) + "): " + what);
    }
    void skip() {
        for (;;) {
            while (p_ < s_.size() && (s_[p_] == ' ' || s_[p_] == '\t' || s_[p_] == '\n' || s_[p_] == '\r')) ++p_;
            if (p_ + 1 < s_.size() && s_[p_] == '/' && s_[p_ + 1] == '/') {
                while (p_ < s_.size() && s_[p_] != '\n') ++p_;
            } else if (p_ + 1 < s_.size() && s_[p_] == '/' && s_[p_ + 1] == '*') {
                p_ += 2;
                while (p_ + 1 < s_.size() && !(s_[p_] == '*' && s_[p_ + 1] == '/')) ++p_;
                if (p_ + 1 >= s_.size()) error("unterminated comment");
                p_ += 2;
            } else {
                return;
            }
        }
    }
    Json value() {
        skip();
        if (p_ >= s_.size()) error("unexpected end");
        char c = s_[p_];
        if (c == '{') return object();
        if (c == '[') return array();
        if (c == '"') return Json::string(string());
        if (s_.compare(p_, 4, "true") == 0) { p_ += 4; return Json::boolean(true); }
        if (s_.compare(p_, 5, "false") == 0) { p_ += 5; return Json::boolean(false); }
        if (s_.compare(p_, 4, "null") == 0) { p_ += 4; return Json(); }
        return number();
    }
    Json number() {
        const char* begin = s_.c_str() + p_;
        char* end = nullptr;
        double v = std::strtod(begin, &end);
        if (end == begin) error("invalid value");
        p_ += size_t(end - begin);
        return Json::number(v);
    }
    std::string string() {
        ++p_;  // opening quote
        std::string out;
        while (p_ < s_.size() && s_[p_] != '"') {
            char c = s_[p_++];
            if (c == '\\') {
                if (p_ >= s_.size()) error("bad escape");
                char e = s_[p_++];
                switch (e) {
                    case 'n': out += '\n'; break;
                    case 't': out += '\t'; break;
                    case 'r': out += '\r'; break;
                    case 'b': out += '\b'; break;
                    case 'f': out += '\f'; break;
                    case 'u': {
                        if (p_ + 4 > s_.size()) error("bad \\u escape");
                        unsigned cp = unsigned(std::strtoul(s_.substr(p_, 4).c_str(), nullptr, 16));
                        p_ += 4;
                        if (cp < 0x80) out += char(cp);
                        else if (cp < 0x800) { out += char(0xC0 | (cp >> 6)); out += char(0x80 | (cp & 0x3F)); }
                        else { out += char(0xE0 | (cp >> 12)); out += char(0x80 | ((cp >> 6) & 0x3F)); out += char(0x80 | (cp & 0x3F)); }
                        break;
                    }
                    default: out += e;
                }
            } else {
                out += c;
            }
        }
        if (p_ >= s_.size()) error("unterminated string");
        ++p_;
        return out;
    }
    Json array() {
        Json j; j.type = Json::Array;
        ++p_;
        for (;;) {
            skip();
            if (p_ < s_.size() && s_[p_] == ']') { ++p_; return j; }
            j.arr.push_back(value());
            skip();
            if (p_ < s_.size() && s_[p_] == ',') { ++p_; continue; }
            if (p_ < s_.size() && s_[p_] == ']') { ++p_; return j; }
            error("expected ',' or ']'");
        }
    }
    Json object() {
        Json j; j.type = Json::Object;
        ++p_;
        for (;;) {
            skip();
            if (p_ < s_.size() && s_[p_] == '}') { ++p_; return j; }
            if (p_ >= s_.size() || s_[p_] != '"') error("expected a key");
            std::string key = string();
            skip();
            if (p_ >= s_.size() || s_[p_] != ':') error("expected ':'");
            ++p_;
            j.obj.emplace_back(std::move(key), value());
            skip();
            if (p_ < s_.size() && s_[p_] == ',') { ++p_; continue; }
            if (p_ < s_.size() && s_[p_] == '}') { ++p_; return j; }
            error("expected ',' or '}'");
        }
    }
};

inline void json_write(const Json& j, std::string& out, int indent = 0) {
    auto pad = [&](int n) { out.append(size_t(n) * 4, ' '); };
    switch (j.type) {
        case Json::Null: out += "null"; break;
        case Json::Bool: out += j.b ? "true" : "false"; break;
        case Json::Number: {
            char buf[40];
            if (j.is_int) std::snprintf(buf, sizeof buf, "%lld", (long long)j.num);
            else if (j.num == std::floor(j.num) && std::fabs(j.num) < 1e15) std::snprintf(buf, sizeof buf, "%.1f", j.num);
            else std::snprintf(buf, sizeof buf, "%.17g", j.num);
            out += buf;
            break;
        }
        case Json::String: {
            out += '"';
            for (char c : j.str) { if (c == '"' || c == '\\') out += '\\'; out += c; }
            out += '"';
            break;
        }
        case Json::Array:
            out += "[";
            for (size_t i = 0; i < j.arr.size(); ++i) {
                out += i ? ",\n" : "\n"; pad(indent + 1);
                json_write(j.arr[i], out, indent + 1);
            }
            out += "\n"; pad(indent); out += "]";
            break;
        case Json::Object:
            out += "{";
            for (size_t i = 0; i < j.obj.size(); ++i) {
                out += i ? ",\n" : "\n"; pad(indent + 1);
                out += '"' + j.obj[i].first + "\": ";
                json_write(j.obj[i].second, out, indent + 1);
            }
            out += "\n"; pad(indent); out += "}";
            break;
    }
}

}  // namespace skyhost
