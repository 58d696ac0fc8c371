// TEST INFRASTRUCTURE ONLY -- jsoncpp is not in this image.  A small functional stand-in for the part of its API that the
// reference sources compiled into oracle/_ref use (Json::Value look-ups and conversions, Json::Reader::parse), so that the
// reference's own loaders (cKinTree::Load, cKinTree::LoadBodyDefs, ...) can read the character files in the pinning tests.
// Accepts the same input the reference's assets use: standard JSON plus `//` comments.
#pragma once
#include <cstdlib>
#include <fstream>
#include <iterator>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

namespace Json {
class Value {
public:
    enum Type { kNull, kBool, kNum, kStr, kArr, kObj };
    Value() : type_(kNull), num_(0), b_(false) {}
    Value(int v) : type_(kNum), num_(v), b_(false) {}
    Value(double v) : type_(kNum), num_(v), b_(false) {}
    Value(bool v) : type_(kBool), num_(v), b_(v) {}
    Value(const char* s) : type_(kStr), num_(0), b_(false), str_(s) {}
    Value(const std::string& s) : type_(kStr), num_(0), b_(false), str_(s) {}

    bool isNull() const { return type_ == kNull; }
    bool isArray() const { return type_ == kArr; }
    bool isObject() const { return type_ == kObj; }
    bool isString() const { return type_ == kStr; }
    bool isNumeric() const { return type_ == kNum || type_ == kBool; }
    bool isBool() const { return type_ == kBool; }
    unsigned size() const { return type_ == kArr ? (unsigned)arr_.size() : (type_ == kObj ? (unsigned)obj_.size() : 0u); }
    double asDouble() const { return type_ == kBool ? (b_ ? 1.0 : 0.0) : num_; }
    float asFloat() const { return (float)asDouble(); }
    int asInt() const { return (int)asDouble(); }
    unsigned asUInt() const { return (unsigned)asDouble(); }
    bool asBool() const { return type_ == kBool ? b_ : num_ != 0; }
    std::string asString() const { return str_; }

    Value operator[](const std::string& k) const { auto it = obj_.find(k); return it == obj_.end() ? Value() : it->second; }
    Value operator[](const char* k) const { return (*this)[std::string(k)]; }
    Value operator[](int i) const { return (type_ == kArr && i >= 0 && i < (int)arr_.size()) ? arr_[i] : Value(); }
    Value operator[](unsigned i) const { return (*this)[(int)i]; }
    Value get(const std::string& k, const Value& dflt) const { auto it = obj_.find(k); return it == obj_.end() ? dflt : it->second; }
    Value get(const char* k, const Value& dflt) const { return get(std::string(k), dflt); }
    Value get(int i, const Value& dflt) const { return (type_ == kArr && i >= 0 && i < (int)arr_.size()) ? arr_[i] : dflt; }
    Value get(unsigned i, const Value& dflt) const { return get((int)i, dflt); }
    bool isMember(const std::string& k) const { return obj_.count(k) != 0; }

private:
    friend class Reader;
    Type type_;
    double num_;
    bool b_;
    std::string str_;
    std::vector<Value> arr_;
    std::map<std::string, Value> obj_;
};

class Reader {
public:
    bool parse(std::istream& in, Value& root) {
        std::stringstream ss;
        ss << in.rdbuf();
        return parse(ss.str(), root);
    }
    bool parse(const std::string& text, Value& root) {
        s_ = &text; p_ = 0; ok_ = true;
        root = value();
        return ok_;
    }

private:
    const std::string* s_ = nullptr;
    size_t p_ = 0;
    bool ok_ = true;
    char cur() const { return p_ < s_->size() ? (*s_)[p_] : '\0'; }
    void ws() {
        for (;;) {
            char c = cur();
            if (c == ' ' || c == '\t' || c == '\n' || c == '\r') ++p_;
            else if (c == '/' && p_ + 1 < s_->size() && (*s_)[p_ + 1] == '/') { while (p_ < s_->size() && (*s_)[p_] != '\n') ++p_; }
            else break;
        }
    }
    std::string str() {
        std::string out;
        ++p_;
        while (p_ < s_->size() && (*s_)[p_] != '"') {
            if ((*s_)[p_] == '\\' && p_ + 1 < s_->size()) { ++p_; char e = (*s_)[p_]; out += (e == 'n' ? '\n' : e == 't' ? '\t' : e); }
            else out += (*s_)[p_];
            ++p_;
        }
        ++p_;
        return out;
    }
    Value value() {
        ws();
        Value v;
        char c = cur();
        if (c == '{') {
            v.type_ = Value::kObj;
            ++p_; ws();
            if (cur() == '}') { ++p_; return v; }
            for (;;) {
                ws();
                if (cur() != '"') { ok_ = false; return v; }
                std::string k = str();
                ws();
                if (cur() != ':') { ok_ = false; return v; }
                ++p_;
                v.obj_[k] = value();
                ws();
                if (cur() == ',') { ++p_; continue; }
                if (cur() == '}') { ++p_; break; }
                ok_ = false; return v;
            }
        } else if (c == '[') {
            v.type_ = Value::kArr;
            ++p_; ws();
            if (cur() == ']') { ++p_; return v; }
            for (;;) {
                v.arr_.push_back(value());
                ws();
                if (cur() == ',') { ++p_; continue; }
                if (cur() == ']') { ++p_; break; }
                ok_ = false; return v;
            }
        } else if (c == '"') {
            v.type_ = Value::kStr; v.str_ = str();
        } else if (!s_->compare(p_, 4, "true")) { v.type_ = Value::kBool; v.b_ = true; v.num_ = 1; p_ += 4; }
        else if (!s_->compare(p_, 5, "false")) { v.type_ = Value::kBool; p_ += 5; }
        else if (!s_->compare(p_, 4, "null")) { p_ += 4; }
        else {
            const char* beg = s_->c_str() + p_;
            char* end = nullptr;
            v.num_ = std::strtod(beg, &end);
            if (end == beg) { ok_ = false; return v; }
            v.type_ = Value::kNum;
            p_ += (size_t)(end - beg);
        }
        return v;
    }
};
}  // namespace Json
