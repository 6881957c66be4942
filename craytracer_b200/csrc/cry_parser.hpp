// .cry scene-description reader: tokenizer and generic raw-value parser.
// Host-side mirror of the reference's src/scene_parser.rs `tokenizer` (:13-253) and `parser`
// (:255-773) modules: same grammar, same error messages and line:column locations, so the host
// tests can restate tests/test_parser.rs.
#pragma once
#include <charconv>
#include <cstdint>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <vector>

namespace cray {

struct Location {  // scene_parser.rs:1-5
    uint32_t line = 0, column = 0;
};

struct ParserError {  // scene_parser.rs:72-90
    std::string message;
    bool has_location = false;
    Location location;
    static ParserError at(const std::string& msg, Location loc) { return {msg, true, loc}; }
    static ParserError nowhere(const std::string& msg) { return {msg, false, {}}; }
};

enum class Tok { Identifier, Number, String, LeftBrace, RightBrace, LeftBracket, RightBracket, LeftParen, RightParen, Comma, Colon, Eof };

// Rust's `{}` for f64: shortest representation that round-trips, no exponent for "ordinary" magnitudes.
inline std::string display_f64(double v) {
    if (v != v) return "NaN";
    char buf[512];
    auto res = std::to_chars(buf, buf + sizeof(buf), v, std::chars_format::fixed);  // shortest fixed form that round-trips
    return std::string(buf, res.ptr);
}

struct Token {
    Tok kind;
    std::string text;  // identifier / string payload
    double number = 0.0;
    Location location;
    std::string display() const {  // impl Display for TokenValue, scene_parser.rs:40-57
        switch (kind) {
            case Tok::Identifier: return "'" + text + "'";
            case Tok::Number: return "'" + display_f64(number) + "'";
            case Tok::String: return "'" + text + "'";
            case Tok::LeftBrace: return "'{'";
            case Tok::RightBrace: return "'}'";
            case Tok::LeftBracket: return "'['";
            case Tok::RightBracket: return "']'";
            case Tok::LeftParen: return "'('";
            case Tok::RightParen: return "')'";
            case Tok::Comma: return "','";
            case Tok::Colon: return "':'";
            default: return "EOF";
        }
    }
};

// Throws ParserError.
std::vector<Token> tokenize(const std::string& input);

struct RawValue;
using RawValuePtr = std::unique_ptr<RawValue>;

struct RawMap {  // RawValueMap scene_parser.rs:411-415 (ordered here; the reference's HashMap order never matters)
    Location location;
    std::vector<std::pair<std::string, RawValuePtr>> entries;
    RawValue* find(const std::string& key);
    bool has(const std::string& key) const;
};

enum class RawKind { Number, String, Vector, Point, Color, Map, TypedMap, Array };

struct RawValue {  // scene_parser.rs:323-333
    RawKind kind = RawKind::Number;
    double number = 0.0;
    std::string string;
    double xyz[3] = {0, 0, 0};
    std::string type_name;  // TypedMap
    RawMap map;             // Map / TypedMap
    std::vector<RawValuePtr> array;
    std::set<std::string> used_keys;  // TypedRawValueMap::used_keys (unused-key warnings, :571-586)
    std::string debug() const;        // approximation of Rust's {:?} used inside conversion error messages
};

// RawValue::from_tokens scene_parser.rs:336-409; throws ParserError
RawValuePtr parse_raw_value(const std::vector<Token>& tokens, size_t& pos);
RawMap parse_raw_map(const std::vector<Token>& tokens, size_t& pos);

std::string tokens_to_json(const std::vector<Token>& tokens);
std::string raw_value_to_json(const RawValue& v);

}  // namespace cray
