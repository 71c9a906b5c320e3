// TEST INFRASTRUCTURE (oracle build shim) — not product code.
//
// A small, self-contained XML DOM that exposes the subset of the tinyxml2 API
// which the UNMODIFIED reference MJCF reader/writer (src/xml/*.cc under
// /root/reference) calls.  tinyxml2 itself is a git-fetched third-party
// dependency of the reference (cmake/MujocoDependencies.cmake) and is absent
// from this image; this header exists only so that the reference's own parser
// can be compiled into oracle/_ref/libmujoco_ref.so and produce the mjModel the
// parity oracle steps.  Written from the public tinyxml2 interface, not copied.
#ifndef ORACLE_SHIM_TINYXML2_H_
#define ORACLE_SHIM_TINYXML2_H_

#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

namespace tinyxml2 {

enum XMLError { XML_SUCCESS = 0, XML_ERROR_PARSING = 1, XML_ERROR_EMPTY_DOCUMENT = 2 };

class XMLDocument;
class XMLElement;
class XMLComment;
class XMLPrinter;

class XMLAttribute {
 public:
  const char* Name() const { return name_.c_str(); }
  const char* Value() const { return value_.c_str(); }
  const XMLAttribute* Next() const { return next_; }

 private:
  friend class XMLElement;
  friend class XMLDocument;
  std::string name_, value_;
  XMLAttribute* next_ = nullptr;
};

class XMLNode {
 public:
  virtual ~XMLNode() {}
  virtual XMLElement* ToElement() { return nullptr; }
  virtual const XMLElement* ToElement() const { return nullptr; }
  virtual XMLComment* ToComment() { return nullptr; }
  virtual const XMLComment* ToComment() const { return nullptr; }

  const char* Value() const { return value_.c_str(); }
  int GetLineNum() const { return line_; }
  XMLDocument* GetDocument() const { return doc_; }
  XMLNode* Parent() const { return parent_; }
  bool NoChildren() const { return first_ == nullptr; }
  XMLNode* FirstChild() const { return first_; }
  XMLNode* NextSibling() const { return next_; }

  XMLElement* FirstChildElement(const char* name = nullptr) const;
  XMLElement* NextSiblingElement(const char* name = nullptr) const;

  XMLNode* InsertEndChild(XMLNode* n) {
    Unlink(n);
    n->parent_ = this;
    n->prev_ = last_;
    n->next_ = nullptr;
    if (last_) last_->next_ = n; else first_ = n;
    last_ = n;
    return n;
  }
  XMLNode* LinkEndChild(XMLNode* n) { return InsertEndChild(n); }
  XMLNode* InsertFirstChild(XMLNode* n) {
    Unlink(n);
    n->parent_ = this;
    n->prev_ = nullptr;
    n->next_ = first_;
    if (first_) first_->prev_ = n; else last_ = n;
    first_ = n;
    return n;
  }
  XMLNode* InsertAfterChild(XMLNode* after, XMLNode* n) {
    if (!after || after->parent_ != this) return nullptr;
    if (after == last_) return InsertEndChild(n);
    Unlink(n);
    n->parent_ = this;
    n->prev_ = after;
    n->next_ = after->next_;
    after->next_->prev_ = n;
    after->next_ = n;
    return n;
  }
  void DeleteChild(XMLNode* n) {
    if (!n || n->parent_ != this) return;
    Unlink(n);  // storage is owned by the document pool; just detach
  }
  virtual XMLNode* ShallowClone(XMLDocument* target) const = 0;
  XMLNode* DeepClone(XMLDocument* target) const {
    XMLNode* c = ShallowClone(target);
    for (XMLNode* k = first_; k; k = k->next_) c->InsertEndChild(k->DeepClone(target));
    return c;
  }

 protected:
  friend class XMLDocument;
  friend class XMLPrinter;
  explicit XMLNode(XMLDocument* d) : doc_(d) {}
  static void Unlink(XMLNode* n) {
    XMLNode* p = n->parent_;
    if (!p) return;
    if (n->prev_) n->prev_->next_ = n->next_; else p->first_ = n->next_;
    if (n->next_) n->next_->prev_ = n->prev_; else p->last_ = n->prev_;
    n->parent_ = n->prev_ = n->next_ = nullptr;
  }
  XMLDocument* doc_;
  XMLNode* parent_ = nullptr;
  XMLNode* first_ = nullptr;
  XMLNode* last_ = nullptr;
  XMLNode* prev_ = nullptr;
  XMLNode* next_ = nullptr;
  std::string value_;
  int line_ = 0;
};

class XMLComment : public XMLNode {
 public:
  XMLComment* ToComment() override { return this; }
  const XMLComment* ToComment() const override { return this; }
  XMLNode* ShallowClone(XMLDocument* target) const override;

 private:
  friend class XMLDocument;
  explicit XMLComment(XMLDocument* d) : XMLNode(d) {}
};

class XMLText : public XMLNode {
 public:
  XMLNode* ShallowClone(XMLDocument* target) const override;

 private:
  friend class XMLDocument;
  explicit XMLText(XMLDocument* d) : XMLNode(d) {}
};

class XMLElement : public XMLNode {
 public:
  XMLElement* ToElement() override { return this; }
  const XMLElement* ToElement() const override { return this; }
  const char* Name() const { return Value(); }
  const XMLAttribute* FirstAttribute() const { return attr_; }
  const XMLAttribute* FindAttribute(const char* name) const {
    for (const XMLAttribute* a = attr_; a; a = a->next_)
      if (a->name_ == name) return a;
    return nullptr;
  }
  const char* Attribute(const char* name, const char* value = nullptr) const {
    const XMLAttribute* a = FindAttribute(name);
    if (!a) return nullptr;
    if (!value || a->value_ == value) return a->Value();
    return nullptr;
  }
  inline void SetAttribute(const char* name, const char* value);
  void SetAttribute(const char* name, int v) { SetAttribute(name, std::to_string(v).c_str()); }
  void SetAttribute(const char* name, unsigned v) { SetAttribute(name, std::to_string(v).c_str()); }
  void SetAttribute(const char* name, int64_t v) { SetAttribute(name, std::to_string(v).c_str()); }
  void SetAttribute(const char* name, uint64_t v) { SetAttribute(name, std::to_string(v).c_str()); }
  void SetAttribute(const char* name, bool v) { SetAttribute(name, v ? "true" : "false"); }
  void SetAttribute(const char* name, double v) {
    char buf[64];
    std::snprintf(buf, sizeof(buf), "%.17g", v);
    SetAttribute(name, buf);
  }
  void SetAttribute(const char* name, float v) {
    char buf[64];
    std::snprintf(buf, sizeof(buf), "%.9g", (double)v);
    SetAttribute(name, buf);
  }
  XMLNode* ShallowClone(XMLDocument* target) const override;

 private:
  friend class XMLDocument;
  friend class XMLPrinter;
  explicit XMLElement(XMLDocument* d) : XMLNode(d) {}
  XMLAttribute* attr_ = nullptr;
};

inline XMLElement* XMLNode::FirstChildElement(const char* name) const {
  for (XMLNode* n = first_; n; n = n->next_) {
    XMLElement* e = n->ToElement();
    if (e && (!name || e->value_ == name)) return e;
  }
  return nullptr;
}
inline XMLElement* XMLNode::NextSiblingElement(const char* name) const {
  for (XMLNode* n = next_; n; n = n->next_) {
    XMLElement* e = n->ToElement();
    if (e && (!name || e->value_ == name)) return e;
  }
  return nullptr;
}

class XMLPrinter {
 public:
  XMLPrinter(FILE* file = nullptr, bool compact = false, int depth = 0)
      : compact_(compact) { (void)file; (void)depth; }
  virtual ~XMLPrinter() {}
  const char* CStr() const { return out_.c_str(); }
  int CStrSize() const { return (int)out_.size() + 1; }
  // NOTE: the reference subclass defines a non-virtual PrintSpace (2 spaces);
  // tinyxml2's is virtual.  Keep it virtual here so the override is picked up.
  virtual void PrintSpace(int depth) {
    for (int i = 0; i < depth; ++i) Write("    ");
  }
  void Write(const char* s) { out_ += s; }

 private:
  friend class XMLDocument;
  static std::string Escape(const std::string& s, bool attr) {
    std::string o;
    for (char c : s) {
      switch (c) {
        case '&': o += "&amp;"; break;
        case '<': o += "&lt;"; break;
        case '>': o += "&gt;"; break;
        case '"': if (attr) o += "&quot;"; else o += c; break;
        default: o += c;
      }
    }
    return o;
  }
  void PrintNode(const XMLNode* n, int depth) {
    if (const XMLElement* e = n->ToElement()) {
      if (!compact_) PrintSpace(depth);
      out_ += "<" + e->value_;
      for (const XMLAttribute* a = e->FirstAttribute(); a; a = a->Next())
        out_ += std::string(" ") + a->Name() + "=\"" + Escape(a->Value(), true) + "\"";
      if (e->NoChildren()) {
        out_ += "/>";
        if (!compact_) out_ += "\n";
        return;
      }
      out_ += ">";
      bool only_text = true;
      for (const XMLNode* k = e->FirstChild(); k; k = k->NextSibling())
        if (k->ToElement() || k->ToComment()) only_text = false;
      if (!only_text && !compact_) out_ += "\n";
      for (const XMLNode* k = e->FirstChild(); k; k = k->NextSibling()) PrintNode(k, depth + 1);
      if (!only_text && !compact_) PrintSpace(depth);
      out_ += "</" + e->value_ + ">";
      if (!compact_) out_ += "\n";
    } else if (n->ToComment()) {
      if (!compact_) PrintSpace(depth);
      out_ += "<!--" + n->value_ + "-->";
      if (!compact_) out_ += "\n";
    } else {
      out_ += Escape(n->value_, false);
    }
  }
  bool compact_;
  std::string out_;
};

class XMLDocument : public XMLNode {
 public:
  XMLDocument() : XMLNode(nullptr) { doc_ = this; }
  ~XMLDocument() override {}
  XMLDocument(const XMLDocument&) = delete;
  XMLDocument& operator=(const XMLDocument&) = delete;

  XMLNode* ShallowClone(XMLDocument*) const override { return nullptr; }

  XMLElement* RootElement() { return FirstChildElement(); }
  bool Error() const { return err_ != XML_SUCCESS; }
  XMLError ErrorID() const { return err_; }
  const char* ErrorStr() const { return errstr_.c_str(); }
  int ErrorLineNum() const { return errline_; }
  void ClearError() { err_ = XML_SUCCESS; errstr_.clear(); }

  XMLElement* NewElement(const char* name) {
    XMLElement* e = new XMLElement(this);
    e->value_ = name;
    pool_.emplace_back(e);
    return e;
  }
  XMLComment* NewComment(const char* text) {
    XMLComment* c = new XMLComment(this);
    c->value_ = text;
    pool_.emplace_back(c);
    return c;
  }
  XMLText* NewText(const char* text) {
    XMLText* t = new XMLText(this);
    t->value_ = text;
    pool_.emplace_back(t);
    return t;
  }
  XMLAttribute* NewAttribute() {
    XMLAttribute* a = new XMLAttribute();
    apool_.emplace_back(a);
    return a;
  }
  void Print(XMLPrinter* p) const {
    for (const XMLNode* k = first_; k; k = k->NextSibling()) p->PrintNode(k, 0);
  }

  XMLError Parse(const char* xml, size_t nbytes = static_cast<size_t>(-1)) {
    ClearError();
    first_ = last_ = nullptr;
    if (!xml) return Fail(XML_ERROR_EMPTY_DOCUMENT, "empty document", 0);
    if (nbytes == static_cast<size_t>(-1)) nbytes = std::strlen(xml);
    s_ = xml;
    n_ = nbytes;
    p_ = 0;
    curline_ = 1;
    if (n_ >= 3 && (unsigned char)s_[0] == 0xEF && (unsigned char)s_[1] == 0xBB &&
        (unsigned char)s_[2] == 0xBF) p_ = 3;
    ParseContent(this);
    if (!Error() && !RootElement()) Fail(XML_ERROR_EMPTY_DOCUMENT, "no root element", curline_);
    return err_;
  }

 private:
  XMLError Fail(XMLError e, const std::string& msg, int line) {
    if (err_ == XML_SUCCESS) {
      err_ = e;
      errline_ = line;
      errstr_ = "Error=XML_ERROR_PARSING ErrorID=" + std::to_string((int)e) +
                " (0x" + std::to_string((int)e) + ") Line number=" + std::to_string(line) + ": " + msg;
    }
    return err_;
  }
  bool Eof() const { return p_ >= n_ || s_[p_] == '\0'; }
  char Cur() const { return Eof() ? '\0' : s_[p_]; }
  void Adv() { if (!Eof()) { if (s_[p_] == '\n') ++curline_; ++p_; } }
  bool StartsWith(const char* lit) const {
    size_t L = std::strlen(lit);
    return p_ + L <= n_ && std::strncmp(s_ + p_, lit, L) == 0;
  }
  void SkipWs() { while (!Eof() && std::strchr(" \t\r\n", Cur())) Adv(); }
  bool SkipUntil(const char* lit, std::string* out) {
    while (!Eof()) {
      if (StartsWith(lit)) { for (size_t i = 0; i < std::strlen(lit); ++i) Adv(); return true; }
      if (out) out->push_back(Cur());
      Adv();
    }
    return false;
  }
  static bool NameChar(char c) {
    return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || (c >= '0' && c <= '9') ||
           c == '_' || c == ':' || c == '-' || c == '.' || ((unsigned char)c >= 128);
  }
  std::string ReadName() {
    std::string s;
    while (!Eof() && NameChar(Cur())) { s.push_back(Cur()); Adv(); }
    return s;
  }
  static void AppendUtf8(std::string& o, unsigned long cp) {
    if (cp < 0x80) o.push_back((char)cp);
    else if (cp < 0x800) { o.push_back((char)(0xC0 | (cp >> 6))); o.push_back((char)(0x80 | (cp & 0x3F))); }
    else if (cp < 0x10000) { o.push_back((char)(0xE0 | (cp >> 12))); o.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); o.push_back((char)(0x80 | (cp & 0x3F))); }
    else { o.push_back((char)(0xF0 | (cp >> 18))); o.push_back((char)(0x80 | ((cp >> 12) & 0x3F))); o.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); o.push_back((char)(0x80 | (cp & 0x3F))); }
  }
  static std::string Unescape(const std::string& in) {
    std::string o;
    for (size_t i = 0; i < in.size(); ++i) {
      if (in[i] != '&') { o.push_back(in[i]); continue; }
      size_t semi = in.find(';', i);
      if (semi == std::string::npos) { o.push_back('&'); continue; }
      std::string ent = in.substr(i + 1, semi - i - 1);
      if (ent == "amp") o.push_back('&');
      else if (ent == "lt") o.push_back('<');
      else if (ent == "gt") o.push_back('>');
      else if (ent == "quot") o.push_back('"');
      else if (ent == "apos") o.push_back('\'');
      else if (!ent.empty() && ent[0] == '#') {
        unsigned long cp = (ent.size() > 1 && (ent[1] == 'x' || ent[1] == 'X'))
                               ? std::strtoul(ent.c_str() + 2, nullptr, 16)
                               : std::strtoul(ent.c_str() + 1, nullptr, 10);
        AppendUtf8(o, cp);
      } else { o.push_back('&'); continue; }
      i = semi;
    }
    return o;
  }
  // parse children of `parent` until its closing tag (or EOF for the document)
  void ParseContent(XMLNode* parent) {
    while (!Error()) {
      // text run
      size_t start = p_;
      int startline = curline_;
      std::string text;
      while (!Eof() && Cur() != '<') { text.push_back(Cur()); Adv(); }
      (void)start;
      if (parent != this) {
        bool allws = true;
        for (char c : text) if (!std::strchr(" \t\r\n", c)) allws = false;
        if (!allws) {
          XMLText* t = NewText(Unescape(text).c_str());
          t->line_ = startline;
          parent->InsertEndChild(t);
        }
      }
      if (Eof()) {
        if (parent != this) Fail(XML_ERROR_PARSING, "unexpected end of document", curline_);
        return;
      }
      int line = curline_;
      if (StartsWith("<!--")) {
        p_ += 4;
        std::string body;
        if (!SkipUntil("-->", &body)) { Fail(XML_ERROR_PARSING, "unterminated comment", line); return; }
        XMLComment* c = NewComment(body.c_str());
        c->line_ = line;
        parent->InsertEndChild(c);
      } else if (StartsWith("<?")) {
        if (!SkipUntil("?>", nullptr)) { Fail(XML_ERROR_PARSING, "unterminated declaration", line); return; }
      } else if (StartsWith("<![CDATA[")) {
        p_ += 9;
        std::string body;
        if (!SkipUntil("]]>", &body)) { Fail(XML_ERROR_PARSING, "unterminated CDATA", line); return; }
        XMLText* t = NewText(body.c_str());
        t->line_ = line;
        parent->InsertEndChild(t);
      } else if (StartsWith("<!")) {
        if (!SkipUntil(">", nullptr)) { Fail(XML_ERROR_PARSING, "unterminated <! section", line); return; }
      } else if (StartsWith("</")) {
        Adv(); Adv();
        std::string name = ReadName();
        SkipWs();
        if (Cur() != '>') { Fail(XML_ERROR_PARSING, "malformed closing tag", line); return; }
        Adv();
        if (parent == this || name != parent->value_) {
          Fail(XML_ERROR_PARSING, "mismatched element: </" + name + ">", line);
        }
        return;
      } else {
        Adv();  // '<'
        std::string name = ReadName();
        if (name.empty()) { Fail(XML_ERROR_PARSING, "malformed element", line); return; }
        XMLElement* e = NewElement(name.c_str());
        e->line_ = line;
        parent->InsertEndChild(e);
        XMLAttribute* tail = nullptr;
        bool selfclose = false;
        for (;;) {
          SkipWs();
          if (Eof()) { Fail(XML_ERROR_PARSING, "unterminated element <" + name + ">", line); return; }
          if (Cur() == '/') {
            Adv();
            if (Cur() != '>') { Fail(XML_ERROR_PARSING, "malformed element <" + name + ">", line); return; }
            Adv();
            selfclose = true;
            break;
          }
          if (Cur() == '>') { Adv(); break; }
          std::string an = ReadName();
          if (an.empty()) { Fail(XML_ERROR_PARSING, "malformed attribute in <" + name + ">", curline_); return; }
          SkipWs();
          if (Cur() != '=') { Fail(XML_ERROR_PARSING, "attribute without value in <" + name + ">", curline_); return; }
          Adv();
          SkipWs();
          char q = Cur();
          if (q != '"' && q != '\'') { Fail(XML_ERROR_PARSING, "unquoted attribute in <" + name + ">", curline_); return; }
          Adv();
          std::string av;
          while (!Eof() && Cur() != q) { av.push_back(Cur()); Adv(); }
          if (Eof()) { Fail(XML_ERROR_PARSING, "unterminated attribute in <" + name + ">", curline_); return; }
          Adv();
          XMLAttribute* a = NewAttribute();
          a->name_ = an;
          a->value_ = Unescape(av);
          if (tail) tail->next_ = a; else e->attr_ = a;
          tail = a;
        }
        if (!selfclose) ParseContent(e);
      }
    }
  }

  std::vector<std::unique_ptr<XMLNode>> pool_;
  std::vector<std::unique_ptr<XMLAttribute>> apool_;
  XMLError err_ = XML_SUCCESS;
  std::string errstr_;
  int errline_ = 0;
  const char* s_ = nullptr;
  size_t n_ = 0, p_ = 0;
  int curline_ = 1;
};

inline void XMLElement::SetAttribute(const char* name, const char* value) {
  XMLAttribute* last = nullptr;
  for (XMLAttribute* a = attr_; a; a = a->next_) {
    if (a->name_ == name) { a->value_ = value; return; }
    last = a;
  }
  XMLAttribute* a = doc_->NewAttribute();
  a->name_ = name;
  a->value_ = value;
  if (last) last->next_ = a; else attr_ = a;
}

inline XMLNode* XMLComment::ShallowClone(XMLDocument* target) const {
  XMLComment* c = (target ? target : doc_)->NewComment(value_.c_str());
  c->line_ = line_;
  return c;
}
inline XMLNode* XMLText::ShallowClone(XMLDocument* target) const {
  XMLText* t = (target ? target : doc_)->NewText(value_.c_str());
  t->line_ = line_;
  return t;
}
inline XMLNode* XMLElement::ShallowClone(XMLDocument* target) const {
  XMLDocument* d = target ? target : doc_;
  XMLElement* e = d->NewElement(value_.c_str());
  e->line_ = line_;
  for (const XMLAttribute* a = attr_; a; a = a->Next()) e->SetAttribute(a->Name(), a->Value());
  return e;
}

}  // namespace tinyxml2

#endif  // ORACLE_SHIM_TINYXML2_H_
