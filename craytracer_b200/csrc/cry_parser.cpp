// See cry_parser.hpp.  Mirrors src/scene_parser.rs:13-773 of the reference.
#include "cry_parser.hpp"

#include <cstdlib>
#include <sstream>

namespace cray {

namespace {

struct Cursor {  // CharsWithLocation scene_parser.rs:92-130
    const std::string& s;
    size_t i = 0;
    Location loc{1, 1};
    bool done() const { return i >= s.size(); }
    char peek() const { return s[i]; }
    char next() {
        char c = s[i++];
        if (c == '\n') { loc.line += 1; loc.column = 1; }
        else if ((static_cast<unsigned char>(c) & 0xC0) != 0x80) loc.column += 1;  // count chars, not UTF-8 continuation bytes
        return c;
    }
};

bool is_digit(char c) { return c >= '0' && c <= '9'; }
bool is_alpha(char c) { return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z'); }

// Rust's `str::parse::<f64>()` restricted to what tokenize_number can produce: [+-]? digits* ('.' digits*)?
bool parse_rust_f64(const std::string& t, double& out) {
    size_t k = 0;
    if (k < t.size() && (t[k] == '+' || t[k] == '-')) ++k;
    size_t digits = 0;
    for (size_t j = k; j < t.size(); ++j)
        if (is_digit(t[j])) ++digits;
    if (digits == 0) return false;  // "", "+", "-", ".", "+." are errors in Rust
    out = std::strtod(t.c_str(), nullptr);  // correctly rounded, like Rust's parser
    return true;
}

Token number_token(Cursor& c) {  // tokenize_number scene_parser.rs:132-159
    std::string number;
    bool has_dot = false;
    Location location = c.loc;
    if (!c.done() && (c.peek() == '+' || c.peek() == '-')) number.push_back(c.next());
    while (!c.done()) {
        char ch = c.peek();
        if (is_digit(ch)) number.push_back(c.next());
        else if (!has_dot && ch == '.') { has_dot = true; number.push_back(c.next()); }
        else break;
    }
    double v;
    if (!parse_rust_f64(number, v)) throw ParserError::at("Cannot parse '" + number + "' as number", location);
    Token t{Tok::Number, "", v, location};
    return t;
}

Token string_token(Cursor& c) {  // tokenize_string scene_parser.rs:161-172
    Location location = c.loc;
    char start = c.next();
    std::string s;
    while (!c.done()) {
        char ch = c.next();
        if (ch == start) return Token{Tok::String, s, 0.0, location};
        s.push_back(ch);
    }
    throw ParserError::at("Unterminated string", location);
}

std::string json_escape(const std::string& s) {
    std::string o;
    for (char ch : s) {
        if (ch == '"' || ch == '\\') { o.push_back('\\'); o.push_back(ch); }
        else if (ch == '\n') o += "\\n";
        else if (ch == '\r') o += "\\r";
        else if (ch == '\t') o += "\\t";
        else o.push_back(ch);
    }
    return o;
}
std::string json_num(double v) {
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof(buf), v);
    return std::string(buf, r.ptr);
}

const Token& expect(const std::vector<Token>& tokens, size_t& pos, Tok kind, const char* shown) {  // expect_token_variant :269-292
    const Token& t = tokens[pos++];
    if (t.kind != kind) throw ParserError::at(std::string("Expected ") + shown + ", got " + t.display(), t.location);
    return t;
}
double expect_number(const std::vector<Token>& tokens, size_t& pos) {  // :294-304
    const Token& t = tokens[pos++];
    if (t.kind != Tok::Number) throw ParserError::at("Expected number, got " + t.display(), t.location);
    return t.number;
}
void parse_triple(const std::vector<Token>& tokens, size_t& pos, double out[3]) {
    expect(tokens, pos, Tok::LeftParen, "'('");
    out[0] = expect_number(tokens, pos);
    expect(tokens, pos, Tok::Comma, "','");
    out[1] = expect_number(tokens, pos);
    expect(tokens, pos, Tok::Comma, "','");
    out[2] = expect_number(tokens, pos);
    expect(tokens, pos, Tok::RightParen, "')'");
}

}  // namespace

std::vector<Token> tokenize(const std::string& input) {  // scene_parser.rs:174-252
    std::vector<Token> tokens;
    Cursor c{input};
    while (!c.done()) {
        char ch = c.peek();
        switch (ch) {
            case ' ': case '\t': case '\n': case '\r': break;
            case '/': {
                c.next();
                if (!c.done() && c.peek() == '/') {
                    c.next();
                    while (!c.done() && c.peek() != '\n') c.next();
                    // the reference then falls through to the unconditional chars.next() below, which eats the '\n'
                } else {
                    throw ParserError::at("Expected a second '/' to start a comment", c.loc);
                }
                break;
            }
            case '{': tokens.push_back({Tok::LeftBrace, "", 0, c.loc}); break;
            case '}': tokens.push_back({Tok::RightBrace, "", 0, c.loc}); break;
            case '[': tokens.push_back({Tok::LeftBracket, "", 0, c.loc}); break;
            case ']': tokens.push_back({Tok::RightBracket, "", 0, c.loc}); break;
            case '(': tokens.push_back({Tok::LeftParen, "", 0, c.loc}); break;
            case ')': tokens.push_back({Tok::RightParen, "", 0, c.loc}); break;
            case ',': tokens.push_back({Tok::Comma, "", 0, c.loc}); break;
            case ':': tokens.push_back({Tok::Colon, "", 0, c.loc}); break;
            case '"': case '\'': tokens.push_back(string_token(c)); continue;
            default:
                if (is_digit(ch) || ch == '+' || ch == '-') { tokens.push_back(number_token(c)); continue; }
                if (is_alpha(ch) || ch == '_') {
                    Location location = c.loc;
                    std::string ident;
                    ident.push_back(c.next());
                    while (!c.done() && (is_alpha(c.peek()) || is_digit(c.peek()) || c.peek() == '_')) ident.push_back(c.next());
                    tokens.push_back({Tok::Identifier, ident, 0, location});
                    continue;
                }
                {
                    // show the whole (possibly multi-byte) character like Rust's `char` would
                    size_t len = 1;
                    unsigned char u = static_cast<unsigned char>(ch);
                    if (u >= 0xF0) len = 4; else if (u >= 0xE0) len = 3; else if (u >= 0xC0) len = 2;
                    throw ParserError::at("Unexpected character: '" + input.substr(c.i, len) + "'", c.loc);
                }
        }
        if (!c.done()) c.next();
    }
    tokens.push_back({Tok::Eof, "", 0, c.loc});
    return tokens;
}

RawValue* RawMap::find(const std::string& key) {
    for (auto& e : entries)
        if (e.first == key) return e.second.get();
    return nullptr;
}
bool RawMap::has(const std::string& key) const {
    for (auto& e : entries)
        if (e.first == key) return true;
    return false;
}

RawMap parse_raw_map(const std::vector<Token>& tokens, size_t& pos) {  // RawValueMap::from_tokens :424-467
    RawMap map;
    const Token& start = expect(tokens, pos, Tok::LeftBrace, "'{'");
    map.location = start.location;
    for (;;) {
        const Token& t = tokens[pos];
        if (t.kind == Tok::Identifier) {
            std::string key = t.text;
            pos++;
            expect(tokens, pos, Tok::Colon, "':'");
            RawValuePtr value = parse_raw_value(tokens, pos);
            if (map.has(key)) throw ParserError::at("Duplicate key " + key, map.location);
            map.entries.emplace_back(key, std::move(value));
        } else {
            break;
        }
        const Token& sep = tokens[pos];
        if (sep.kind == Tok::Comma) pos++;
        else break;
    }
    expect(tokens, pos, Tok::RightBrace, "'}'");
    return map;
}

RawValuePtr parse_raw_value(const std::vector<Token>& tokens, size_t& pos) {  // RawValue::from_tokens :336-409
    const Token& token = tokens[pos];
    auto v = std::make_unique<RawValue>();
    switch (token.kind) {
        case Tok::Number:
            pos++;
            v->kind = RawKind::Number;
            v->number = token.number;
            return v;
        case Tok::String:
            pos++;
            v->kind = RawKind::String;
            v->string = token.text;
            return v;
        case Tok::Identifier: {
            pos++;
            const Token& opener = tokens[pos];
            if (opener.kind == Tok::LeftParen) {
                if (token.text == "Vector" || token.text == "Point" || token.text == "Color") {
                    v->kind = token.text == "Vector" ? RawKind::Vector : (token.text == "Point" ? RawKind::Point : RawKind::Color);
                    parse_triple(tokens, pos, v->xyz);
                    return v;
                }
                // TypedRawValueMap::from_tokens (:527-538) expects an identifier where the '(' is
                throw ParserError::at("Expected identifier, got " + opener.display(), opener.location);
            }
            if (opener.kind == Tok::LeftBrace) {
                v->kind = RawKind::TypedMap;
                v->type_name = token.text;
                v->map = parse_raw_map(tokens, pos);
                return v;
            }
            throw ParserError::at("Expected '(' or '{', got " + opener.display(), opener.location);
        }
        case Tok::LeftBrace:
            v->kind = RawKind::Map;
            v->map = parse_raw_map(tokens, pos);
            return v;
        case Tok::LeftBracket: {  // RawValueArray::from_tokens :593-628
            v->kind = RawKind::Array;
            expect(tokens, pos, Tok::LeftBracket, "'['");
            for (;;) {
                if (tokens[pos].kind == Tok::RightBracket) break;
                v->array.push_back(parse_raw_value(tokens, pos));
                if (tokens[pos].kind == Tok::Comma) pos++;
                else break;
            }
            expect(tokens, pos, Tok::RightBracket, "']'");
            return v;
        }
        default:
            throw ParserError::at("Expected a raw value. Got " + token.display(), token.location);
    }
}

std::string RawValue::debug() const {
    switch (kind) {
        case RawKind::Number: return "Number(" + json_num(number) + (number == (double)(long long)number ? ".0" : "") + ")";
        case RawKind::String: return "String(\"" + string + "\")";
        case RawKind::Vector: return "Vector(..)";
        case RawKind::Point: return "Point(..)";
        case RawKind::Color: return "Color(..)";
        case RawKind::Map: return "Map(..)";
        case RawKind::TypedMap: return "TypedMap(" + type_name + " ..)";
        default: return "Array(..)";
    }
}

std::string tokens_to_json(const std::vector<Token>& tokens) {
    static const char* names[] = {"Identifier", "Number", "String", "LeftBrace", "RightBrace", "LeftBracket", "RightBracket", "LeftParen", "RightParen", "Comma", "Colon", "Eof"};
    std::ostringstream o;
    o << "[";
    for (size_t i = 0; i < tokens.size(); ++i) {
        const Token& t = tokens[i];
        if (i) o << ",";
        o << "{\"kind\":\"" << names[(int)t.kind] << "\",\"line\":" << t.location.line << ",\"column\":" << t.location.column;
        if (t.kind == Tok::Number) o << ",\"value\":" << json_num(t.number);
        if (t.kind == Tok::Identifier || t.kind == Tok::String) o << ",\"value\":\"" << json_escape(t.text) << "\"";
        o << "}";
    }
    o << "]";
    return o.str();
}

static void map_to_json(const RawMap& m, std::ostringstream& o) {
    o << "{\"line\":" << m.location.line << ",\"column\":" << m.location.column << ",\"entries\":{";
    for (size_t i = 0; i < m.entries.size(); ++i) {
        if (i) o << ",";
        o << "\"" << json_escape(m.entries[i].first) << "\":" << raw_value_to_json(*m.entries[i].second);
    }
    o << "}}";
}

std::string raw_value_to_json(const RawValue& v) {
    std::ostringstream o;
    switch (v.kind) {
        case RawKind::Number: o << "{\"t\":\"Number\",\"v\":" << json_num(v.number) << "}"; break;
        case RawKind::String: o << "{\"t\":\"String\",\"v\":\"" << json_escape(v.string) << "\"}"; break;
        case RawKind::Vector:
        case RawKind::Point:
        case RawKind::Color:
            o << "{\"t\":\"" << (v.kind == RawKind::Vector ? "Vector" : (v.kind == RawKind::Point ? "Point" : "Color")) << "\",\"v\":[" << json_num(v.xyz[0]) << ","
              << json_num(v.xyz[1]) << "," << json_num(v.xyz[2]) << "]}";
            break;
        case RawKind::Map:
            o << "{\"t\":\"Map\",\"v\":";
            map_to_json(v.map, o);
            o << "}";
            break;
        case RawKind::TypedMap:
            o << "{\"t\":\"TypedMap\",\"name\":\"" << json_escape(v.type_name) << "\",\"v\":";
            map_to_json(v.map, o);
            o << "}";
            break;
        case RawKind::Array:
            o << "{\"t\":\"Array\",\"v\":[";
            for (size_t i = 0; i < v.array.size(); ++i) {
                if (i) o << ",";
                o << raw_value_to_json(*v.array[i]);
            }
            o << "]}";
            break;
    }
    return o.str();
}

}  // namespace cray
