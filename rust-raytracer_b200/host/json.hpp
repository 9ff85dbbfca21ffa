// Minimal JSON reader (objects, arrays, numbers, strings, true/false/null) for the scene files.
// Mirrors what serde_json accepts for the reference's Config schema (config.rs:66-75); unknown keys are ignored by the caller.
#pragma once
#include <cmath>
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
namespace rthost {
struct Json {
    enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
    bool b = false; double num = 0; std::string str;
    std::vector<Json> arr; std::vector<std::pair<std::string, Json>> obj;   // insertion order kept (externally tagged enums)
    const Json* find(const std::string& k) const { for (auto& kv : obj) if (kv.first == k) return &kv.second; return nullptr; }
    const Json& at(const std::string& k) const { const Json* j = find(k); if (!j) throw std::runtime_error("missing field `" + k + "`"); return *j; }
    double number() const { if (kind != Num) throw std::runtime_error("number expected"); return num; }
    const std::string& string() const { if (kind != Str) throw std::runtime_error("string expected"); return str; }
};
class JsonParser {
    const char* p; const char* e;
    int depth = 0;                              // serde_json's default recursion limit is 128 ("recursion limit exceeded")
    struct Nest { JsonParser& q; explicit Nest(JsonParser& q_) : q(q_) { if (++q.depth > 128) q.err("recursion limit exceeded"); } ~Nest() { --q.depth; } };
    void ws() { while (p < e && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p; }
    [[noreturn]] void err(const char* m) { throw std::runtime_error(std::string("json: ") + m); }
    Json value() {
        ws(); if (p >= e) err("unexpected end");
        Nest nest(*this);
        Json j;
        if (*p == '{') { ++p; j.kind = Json::Obj; ws(); if (p < e && *p == '}') { ++p; return j; }
            for (;;) { ws(); if (p >= e || *p != '"') err("key expected"); std::string k = str(); ws(); if (p >= e || *p != ':') err("':' expected"); ++p;
                j.obj.emplace_back(k, value()); ws(); if (p < e && *p == ',') { ++p; continue; } if (p < e && *p == '}') { ++p; return j; } err("',' or '}' expected"); } }
        if (*p == '[') { ++p; j.kind = Json::Arr; ws(); if (p < e && *p == ']') { ++p; return j; }
            for (;;) { j.arr.push_back(value()); ws(); if (p < e && *p == ',') { ++p; continue; } if (p < e && *p == ']') { ++p; return j; } err("',' or ']' expected"); } }
        if (*p == '"') { j.kind = Json::Str; j.str = str(); return j; }
        if (e - p >= 4 && std::string(p, 4) == "true") { p += 4; j.kind = Json::Bool; j.b = true; return j; }
        if (e - p >= 5 && std::string(p, 5) == "false") { p += 5; j.kind = Json::Bool; return j; }
        if (e - p >= 4 && std::string(p, 4) == "null") { p += 4; return j; }
        char* end = nullptr; j.num = std::strtod(p, &end); if (end == p) err("value expected"); p = end; j.kind = Json::Num; return j;
    }
    std::string str() { std::string s; ++p;
        while (p < e && *p != '"') { if (*p == '\\' && p + 1 < e) { ++p; switch (*p) { case 'n': s += '\n'; break; case 't': s += '\t'; break; case 'r': s += '\r'; break;
            case 'b': s += '\b'; break; case 'f': s += '\f'; break; case 'u': { if (e - p < 5) err("bad \\u"); unsigned cp = (unsigned)std::strtoul(std::string(p + 1, 4).c_str(), nullptr, 16); p += 4;
                if (cp < 0x80) s += (char)cp; else if (cp < 0x800) { s += (char)(0xC0 | (cp >> 6)); s += (char)(0x80 | (cp & 63)); } else { s += (char)(0xE0 | (cp >> 12)); s += (char)(0x80 | ((cp >> 6) & 63)); s += (char)(0x80 | (cp & 63)); } break; }
            default: s += *p; } ++p; } else s += *p++; }
        if (p >= e) { err("unterminated string"); }
        ++p;
        return s; }
public:
    static Json parse(const std::string& text) { JsonParser q; q.p = text.data(); q.e = q.p + text.size(); Json j = q.value(); q.ws(); if (q.p != q.e) q.err("trailing characters"); return j; }
};
}  // namespace rthost
