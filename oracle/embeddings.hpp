// ORACLE — TEST INFRASTRUCTURE ONLY (see ojson.hpp header).
//
// CPU restatement of the /v1/embeddings request path:
//   EmbeddingsEndpointSpec.ParseBody            internal/endpointspec/endpointspec.go:231-240
//   EmbeddingRequest / input union              internal/apischema/openai/openai.go:316-375,1612-1665; internal/apischema/openai/union.go:71-185
//   OpenAI→OpenAI RequestBody                   internal/translator/openai_embeddings.go:38-59
//   OpenAI→Azure OpenAI RequestBody             internal/translator/openai_azureopenai_embeddings.go:36-61
//   OpenAI→GCP Vertex AI RequestBody            internal/translator/openai_gcpvertexai_embeddings.go:46-180; internal/apischema/gcp/gcp.go:45-76
// Pinned by the byte-exact goldens tests/data-plane/testupstream_test.go:782,796,810,824,838,852 (Vertex) and the
// passthrough cases of the same table.  Integer-token inputs ([]int64, [][]int64) are type-checked like the reference does
// and pass through the OpenAI translator untouched; the Vertex translator rejects them ("unsupported input type").
#pragma once
#include "translate.hpp"

namespace oracle {

struct EmbItem { bool content_is_array = false; std::vector<std::string> content; std::string task_type, title; };
struct EmbReq {
  enum Kind { NONE, STR, STRS, ITEM, ITEMS, INTS, INT_ARRAYS } kind = NONE;
  std::string str; std::vector<std::string> strs; std::vector<EmbItem> items;
  std::string model; std::optional<int64_t> dimensions; bool has_vendor = false, auto_truncate = false; std::string task_type;
};

inline bool emb_strings(const Value& a, std::vector<std::string>& out) {  // json.Unmarshal into []string: null elements stay ""
  for (auto& e : a.arr) { if (e.is_null()) out.emplace_back(); else if (e.is_str()) out.push_back(e.s); else return false; }
  return true;
}
inline Error emb_item(const Value& v, EmbItem& it) {  // EmbeddingInputItem + EmbeddingContent.UnmarshalJSON
  if (v.is_null()) return {};
  if (!v.is_obj()) return bad("json: cannot unmarshal into Go value of type openai.EmbeddingInputItem");
  if (const Value* c = v.get("content")) {
    if (c->is_null() || c->is_str()) { it.content_is_array = false; it.content.assign(1, c->is_str() ? c->s : std::string()); }
    else if (c->is_arr()) { it.content_is_array = true; if (!emb_strings(*c, it.content)) return bad("content must be string or array of strings"); }
    else return bad("content must be string or array of strings");
  }
  std::optional<std::string> s;
  if (!opt_str(v.get("task_type"), s)) return bad("json: cannot unmarshal into Go struct field EmbeddingInputItem.task_type of type string");
  if (s) it.task_type = *s;
  s.reset();
  if (!opt_str(v.get("title"), s)) return bad("json: cannot unmarshal into Go struct field EmbeddingInputItem.title of type string");
  if (s) it.title = *s;
  return {};
}
inline bool emb_item_empty(const EmbItem& it) { return it.content.empty() || (!it.content_is_array && it.content[0].empty()); }

inline Error parse_embedding_request(const Value& root, EmbReq& r) {
  if (root.is_null()) return {};
  if (!root.is_obj()) return bad("json: cannot unmarshal non-object into Go value of type openai.EmbeddingRequest");
  if (const Value* in = root.get("input")) {  // unmarshalJSONEmbeddingInput (union.go:71-147)
    if (in->is_str()) { r.kind = EmbReq::STR; r.str = in->s; }
    else if (in->is_obj()) {
      EmbItem it; if (auto e = emb_item(*in, it)) return bad("cannot unmarshal input as EmbeddingInputItem: " + e.msg);
      if (emb_item_empty(it)) return bad("invalid input type (must be string, object, or array)");
      r.kind = EmbReq::ITEM; r.items.push_back(it);
    } else if (in->is_arr()) {
      if (in->arr.empty()) r.kind = EmbReq::STRS;
      else {
        const Value& f = in->arr[0];
        if (f.is_str()) { r.kind = EmbReq::STRS; if (!emb_strings(*in, r.strs)) return bad("cannot unmarshal input as []string"); }
        else if (f.is_obj()) {
          r.kind = EmbReq::ITEMS;
          for (auto& e : in->arr) { EmbItem it; if (auto er = emb_item(e, it)) return bad("cannot unmarshal input as []EmbeddingInputItem: " + er.msg); r.items.push_back(it); }
          for (auto& it : r.items) if (emb_item_empty(it)) return bad("invalid input array element");
        } else if (f.is_arr()) {
          r.kind = EmbReq::INT_ARRAYS;
          for (auto& e : in->arr) { if (e.is_null()) continue; if (!e.is_arr()) return bad("cannot unmarshal input as [][]int64"); for (auto& x : e.arr) { int64_t q; if (!int_field(&x, q)) return bad("cannot unmarshal input as [][]int64"); } }
        } else if (f.is_num()) {
          r.kind = EmbReq::INTS;
          for (auto& x : in->arr) { int64_t q; if (!int_field(&x, q)) return bad("cannot unmarshal input as []int64"); }
        } else return bad("invalid input array element");
      }
    } else return bad("invalid input type (must be string, object, or array)");
  }
  std::optional<std::string> s;
  if (!opt_str(root.get("model"), s)) return bad("json: cannot unmarshal into Go struct field EmbeddingRequest.model of type string");
  if (s) r.model = *s;
  if (!str_or_null(root.get("encoding_format")) || !str_or_null(root.get("user"))) return bad("json: cannot unmarshal into Go struct field of type string");
  if (const Value* d = root.get("dimensions"); d && !d->is_null()) { int64_t q; if (!int_field(d, q)) return bad("json: cannot unmarshal into Go struct field EmbeddingRequest.dimensions of type int"); r.dimensions = q; }
  if (const Value* a = root.get("auto_truncate")) { r.has_vendor = true; if (!a->is_null()) { if (!a->is_bool()) return bad("json: cannot unmarshal into Go struct field EmbeddingRequest.auto_truncate of type bool"); r.auto_truncate = a->t == oj::T::True; } }
  if (const Value* t = root.get("task_type")) { r.has_vendor = true; s.reset(); if (!opt_str(t, s)) return bad("json: cannot unmarshal into Go struct field EmbeddingRequest.task_type of type string"); if (s) r.task_type = *s; }
  return {};
}

inline void emb_instances_from_item(const EmbItem& it, std::vector<EmbItem>& out) {  // createInstancesFromEmbeddingInputItem
  for (auto& c : it.content) {
    EmbItem o; o.content.assign(1, c); o.task_type = it.task_type;
    if (it.task_type == "RETRIEVAL_DOCUMENT" && !it.title.empty()) o.title = it.title;
    out.push_back(o);
  }
}

// schema: SCHEMA_OPENAI, SCHEMA_AZURE_OPENAI, SCHEMA_AWS_BEDROCK (Titan) or SCHEMA_GCP_VERTEX.  prefix: OpenAI path prefix ("v1").
inline TranslateResult embeddings_translate(int schema, std::string_view body, const std::string& model_override, const std::string& prefix, bool force) {
  TranslateResult res;
  Value root; std::string perr;
  if (!oj::parse(body, root, perr)) { res.err = bad("malformed request: failed to parse JSON for /v1/embeddings: " + perr); return res; }
  EmbReq r;
  if (auto e = parse_embedding_request(root, r)) { res.err = bad("malformed request: failed to parse JSON for /v1/embeddings: " + e.msg); return res; }
  res.model = r.model; res.request_model = model_override.empty() ? r.model : model_override;
  if (schema == SCHEMA_OPENAI || schema == SCHEMA_AZURE_OPENAI) {
    std::string nb; bool has = false;
    if (!model_override.empty()) { std::string lit; sjson_stringify(lit, model_override); nb = sjson_set_raw(body, root, "model", "", lit); has = true; }
    std::string path = "/";
    if (schema == SCHEMA_AZURE_OPENAI) path = "/openai/deployments/" + res.request_model + "/embeddings?api-version=" + prefix;  // openai_azureopenai_embeddings.go:36-61 (prefix carries the api-version)
    else {
      std::string pfx = prefix; while (!pfx.empty() && pfx.front() == '/') pfx.erase(0, 1); while (!pfx.empty() && pfx.back() == '/') pfx.pop_back();
      if (!pfx.empty()) path += pfx + "/"; path += "embeddings";
    }
    res.headers.push_back({":path", path});
    if (force && (!has || nb.empty())) { nb.assign(body); has = true; }
    if (has && !nb.empty()) { res.body_kind = BYTES; res.body = nb; res.headers.push_back({"content-length", std::to_string(nb.size())}); }
    return res;
  }
  if (schema == SCHEMA_AWS_BEDROCK) {
    // Amazon Titan Embed Text through InvokeModel (openai_awsbedrock_embeddings.go:35-74): one input text, optional dimensions
    std::string text;
    if (r.kind == EmbReq::STR) text = r.str;
    else if (r.kind == EmbReq::STRS) {
      if (r.strs.size() != 1) { res.err = invalid("invalid request body: AWS Bedrock Titan does not support batch embeddings (got " + std::to_string(r.strs.size()) + " inputs)"); return res; }
      text = r.strs[0];
    } else { res.err = invalid("invalid request body: unsupported input type"); return res; }
    std::string o = "{\"inputText\":"; oj::enc_str(o, text);
    if (r.dimensions) o += ",\"dimensions\":" + std::to_string(*r.dimensions);   // *int, omitempty: forwarded whenever present, zero and negatives included
    o.push_back('}');
    res.body_kind = BYTES; res.body = o;
    res.headers.push_back({":path", "/model/" + path_escape(res.request_model) + "/invoke"});
    res.headers.push_back({"content-length", std::to_string(o.size())});
    return res;
  }
  if (schema != SCHEMA_GCP_VERTEX) { res.err = Error{DECLINED, "schema not restated yet"}; return res; }
  std::vector<EmbItem> inst;
  switch (r.kind) {
    case EmbReq::STR: { EmbItem o; o.content.assign(1, r.str); inst.push_back(o); break; }
    case EmbReq::STRS: for (auto& s : r.strs) { EmbItem o; o.content.assign(1, s); inst.push_back(o); } break;
    case EmbReq::ITEM: case EmbReq::ITEMS: for (auto& it : r.items) emb_instances_from_item(it, inst); break;
    default: res.err = internal("error converting EmbeddingRequest: unsupported input type for embedding"); return res;
  }
  if (r.has_vendor && !r.task_type.empty()) for (auto& o : inst) o.task_type = r.task_type;
  std::string o = "{\"instances\":";
  if (inst.empty()) o += "null";
  else {
    o.push_back('[');
    for (size_t i = 0; i < inst.size(); i++) {
      if (i) o.push_back(',');
      o += "{\"content\":"; oj::enc_str(o, inst[i].content[0]);
      if (!inst[i].task_type.empty()) { o += ",\"task_type\":"; oj::enc_str(o, inst[i].task_type); }
      if (!inst[i].title.empty()) { o += ",\"title\":"; oj::enc_str(o, inst[i].title); }
      o.push_back('}');
    }
    o.push_back(']');
  }
  o += ",\"parameters\":{"; bool f = true;
  if (r.has_vendor && r.auto_truncate) { o += "\"auto_truncate\":true"; f = false; }
  if (r.dimensions && *r.dimensions > 0) { if (!f) o.push_back(','); o += "\"outputDimensionality\":" + std::to_string(*r.dimensions); }
  o += "}}";
  res.body_kind = BYTES; res.body = o;
  res.headers.push_back({":path", "publishers/google/models/" + res.request_model + ":predict"});
  res.headers.push_back({"content-length", std::to_string(o.size())});
  return res;
}

}  // namespace oracle
