// json_min.hpp — a small recursive-descent JSON reader (objects, arrays, strings, numbers, true/false/null), enough for
// opt.json (RenderOptions) and transforms_*.json (blender poses).  Replaces the reference's use of nlohmann::json
// (renderer/3rdparty/json.hpp) on this path.  `at(key)` throws like nlohmann's `.at()` so a missing RenderOptions key
// is an error, as in the reference (render_options.hpp:61-77).
#pragma once
#include <cmath>
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace rtohost {

struct Json {
    enum Type { Null, Bool, Number, String, Array, Object } type = Null;
    bool b = false;
    double num = 0.0;
    std::string str;
    std::vector<Json> arr;
    std::map<std::string, Json> obj;

    const Json& at(const std::string& k) const {
        if (type != Object) throw std::runtime_error("json: not an object");
        auto it = obj.find(k);
        if (it == obj.end()) throw std::out_of_range("json: key '" + k + "' not found");
        return it->second;
    }
    bool contains(const std::string& k) const { return type == Object && obj.count(k); }
    const Json& operator[](size_t i) const {
        if (type != Array || i >= arr.size()) throw std::out_of_range("json: array index out of range");
        return arr[i];
    }
    size_t size() const { return type == Array ? arr.size() : obj.size(); }
    double as_number() const {
        if (type != Number) throw std::runtime_error("json: not a number");
        return num;
    }
    bool as_bool() const {
        if (type != Bool) throw std::runtime_error("json: not a boolean");
        return b;
    }

    static Json parse(const std::string& text) {
        size_t i = 0;
        Json v = parse_value(text, i);
        skip_ws(text, i);
        if (i != text.size()) throw std::runtime_error("json: trailing characters");
        return v;
    }

   private:
    static void skip_ws(const std::string& s, size_t& i) {
        while (i < s.size() && (s[i] == ' ' || s[i] == '\n' || s[i] == '\t' || s[i] == '\r')) ++i;
    }
    static Json parse_value(const std::string& s, size_t& i) {
        skip_ws(s, i);
        if (i >= s.size()) throw std::runtime_error("json: unexpected end");
        Json v;
        const char c = s[i];
        if (c == '{') {
            v.type = Object;
            ++i;
            skip_ws(s, i);
            if (i < s.size() && s[i] == '}') { ++i; return v; }
            for (;;) {
                skip_ws(s, i);
                Json k = parse_value(s, i);
                if (k.type != String) throw std::runtime_error("json: object key must be a string");
                skip_ws(s, i);
                if (i >= s.size() || s[i] != ':') throw std::runtime_error("json: expected ':'");
                ++i;
                v.obj[k.str] = parse_value(s, i);
                skip_ws(s, i);
                if (i < s.size() && s[i] == ',') { ++i; continue; }
                if (i < s.size() && s[i] == '}') { ++i; break; }
                throw std::runtime_error("json: expected ',' or '}'");
            }
        } else if (c == '[') {
            v.type = Array;
            ++i;
            skip_ws(s, i);
            if (i < s.size() && s[i] == ']') { ++i; return v; }
            for (;;) {
                v.arr.push_back(parse_value(s, i));
                skip_ws(s, i);
                if (i < s.size() && s[i] == ',') { ++i; continue; }
                if (i < s.size() && s[i] == ']') { ++i; break; }
                throw std::runtime_error("json: expected ',' or ']'");
            }
        } else if (c == '"') {
            v.type = String;
            ++i;
            while (i < s.size() && s[i] != '"') {
                if (s[i] == '\\' && i + 1 < s.size()) {
                    ++i;
                    switch (s[i]) {
                        case 'n': v.str.push_back('\n'); break;
                        case 't': v.str.push_back('\t'); break;
                        case 'r': v.str.push_back('\r'); break;
                        case 'u': v.str.push_back('?'); i += 4; break;
                        default: v.str.push_back(s[i]);
                    }
                    ++i;
                } else {
                    v.str.push_back(s[i++]);
                }
            }
            if (i >= s.size()) throw std::runtime_error("json: unterminated string");
            ++i;
        } else if (s.compare(i, 4, "true") == 0) {
            v.type = Bool; v.b = true; i += 4;
        } else if (s.compare(i, 5, "false") == 0) {
            v.type = Bool; v.b = false; i += 5;
        } else if (s.compare(i, 4, "null") == 0) {
            v.type = Null; i += 4;
        } else {
            char* end = nullptr;
            v.num = strtod(s.c_str() + i, &end);
            if (end == s.c_str() + i) throw std::runtime_error("json: bad token");
            v.type = Number;
            i = (size_t)(end - s.c_str());
        }
        return v;
    }
};

}  // namespace rtohost
