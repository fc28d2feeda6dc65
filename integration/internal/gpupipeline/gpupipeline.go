// Package gpupipeline binds libaigw_b200.so (include/aigw_b200.h) into the extproc handlers of envoyproxy/ai-gateway.
//
// This file is the source a maintainer adds to the reference tree as internal/gpupipeline/gpupipeline.go (INTEGRATION.md says
// where it is called from).  It is not built in this repository: the image has no Go toolchain.  aigw_b200/capi.py makes the
// same C calls over ctypes and is what the tests and the bench run.
package gpupipeline

/*
#cgo CFLAGS: -I${SRCDIR}/../../third_party/aigw_b200/include
#cgo LDFLAGS: -L${SRCDIR}/../../third_party/aigw_b200 -laigw_b200
#include <stdlib.h>
#include "aigw_b200.h"
*/
import "C"

import (
	"errors"
	"unsafe"
)

// ErrDeclined: the body is outside the GPU fast path; the caller runs the stock translator for this request.
var ErrDeclined = errors.New("gpupipeline: declined")

// Pipeline is one GPU: one aigw_ctx (batch calls, one at a time) and one batcher (single-request calls, concurrent).
type Pipeline struct {
	ctx     *C.aigw_ctx
	batcher *C.aigw_batcher
}

// Backend indexes a backend registered with the batcher (one per filterapi.Backend schema / override combination).
type Backend int

func New(device int, maxBatch, windowMicros uint32, first C.aigw_backend_cfg) (*Pipeline, error) {
	p := &Pipeline{}
	if rc := C.aigw_init(C.int(device), &p.ctx); rc != 0 {
		return nil, errors.New("aigw_init failed") // no CPU path inside the library: surface the error, keep the stock handlers
	}
	C.aigw_bind_numa(C.int(device)) // pinned arenas on the GPU's socket
	if rc := C.aigw_batcher_start(p.ctx, &first, C.uint32_t(maxBatch), C.uint32_t(windowMicros), &p.batcher); rc != 0 {
		C.aigw_destroy(p.ctx)
		return nil, errors.New("aigw_batcher_start failed")
	}
	return p, nil
}

// AddBackend registers another backend configuration (schema, model override, prefix, api-version) with the batcher.
func (p *Pipeline) AddBackend(cfg C.aigw_backend_cfg) (Backend, error) {
	idx := C.aigw_batcher_add_backend(p.batcher, &cfg)
	if idx < 0 {
		return 0, errors.New("aigw_batcher_add_backend failed")
	}
	return Backend(idx), nil
}

// Result of one request: what ParseBody + Translator.RequestBody return today (translator.go:41-76).
type Result struct {
	Status int    // AIGW_OK, AIGW_DECLINED, AIGW_MALFORMED_400, AIGW_INVALID_422, AIGW_INTERNAL
	Reason int    // aigw_reason (diagnostics; the shim builds the user-facing message of 400 / 422 from it)
	Path   []byte // ":path" header value
	Body   []byte // translated body, nil when Unchanged
	Model  []byte // ParseBody's originalModel: a slice of the caller's body
	Stream bool
	Unchanged bool // AIGW_BODY_UNCHANGED: forward the original body (no body mutation)
}

// Translate is the per-request call: upstreamProcessor.ProcessRequestHeaders (processor_impl.go:307-398) calls it where it
// calls u.translator.RequestBody today.  It blocks this goroutine's OS thread until the request's batch has run (window +
// one fused kernel, ~0.1-0.6 ms under load).  scratch must hold the translated record (len(body)*5/4 + 1 KiB is always enough).
func (p *Pipeline) Translate(be Backend, body, scratch []byte) (Result, error) {
	var res C.aigw_doc_result
	rc := C.aigw_batcher_translate_to(p.batcher, C.int(be), (*C.uint8_t)(unsafe.Pointer(&body[0])), C.uint32_t(len(body)),
		(*C.uint8_t)(unsafe.Pointer(&scratch[0])), C.uint32_t(len(scratch)), &res)
	if rc == -4 { // scratch too small: res.body_len says how much is needed
		return Result{Status: int(C.AIGW_DECLINED)}, ErrDeclined
	}
	if rc != 0 {
		return Result{Status: int(C.AIGW_DECLINED)}, ErrDeclined
	}
	r := Result{Status: int(res.status), Reason: int(res.reason), Stream: res.flags&1 != 0, Unchanged: res.body_kind == C.AIGW_BODY_UNCHANGED}
	if res.status == C.AIGW_DECLINED {
		return r, ErrDeclined
	}
	if res.status != C.AIGW_OK {
		return r, nil // 400 / 422 / internal: internalapi.ErrMalformedRequest / ErrInvalidRequestBody, as today
	}
	pl, bl := int(res.path_len), int(res.body_len)
	r.Path = scratch[:pl]
	if !r.Unchanged {
		r.Body = scratch[pl : pl+bl]
	}
	r.Model = body[int(res.model_off) : int(res.model_off)+int(res.model_len)]
	return r, nil
}

// ---- response streams: Translator.ResponseBody(hdrs, body, endOfStream), one call per upstream chunk -------------------

// Stream is one response stream's device-resident state (parser state + the undecoded tail of the bytes fed so far).
type Stream struct {
	p *Pipeline
	h C.uint64_t
}

// OpenStream is called where the translator for a streaming response is created (after ResponseHeaders).
// kind: AIGW_STREAM_OPENAI / AWS_BEDROCK / GCP_ANTHROPIC / GCP_GEMINI / GCP_GEMINI_BUFFERED.
func (p *Pipeline) OpenStream(kind int, created int64, requestModel, responseID string) (*Stream, error) {
	cfg := C.aigw_stream_cfg{kind: C.int32_t(kind), created: C.int64_t(created)}
	rm, rid := C.CString(requestModel), C.CString(responseID)
	defer C.free(unsafe.Pointer(rm))
	defer C.free(unsafe.Pointer(rid))
	cfg.request_model, cfg.response_id = rm, rid
	var h C.uint64_t
	if rc := C.aigw_stream_open(p.ctx, &cfg, &h); rc != 0 {
		return nil, errors.New("aigw_stream_open failed")
	}
	return &Stream{p: p, h: h}, nil
}

// ChunkResult mirrors what ResponseBody returns for one chunk: the body mutation (or none), the token usage so far, the
// response model.
type ChunkResult struct {
	Status    int
	Body      []byte // replacement bytes for this chunk (nil when Unchanged)
	Unchanged bool
	Usage     C.aigw_usage
	Model     []byte
}

// Chunk feeds one upstream chunk.  The per-GPU stream goroutine normally batches the chunks of all streams that arrived in a
// window into ONE aigw_stream_chunks call (see INTEGRATION.md); this single-chunk form is the same call with n = 1.
func (s *Stream) Chunk(b []byte, endOfStream bool, out []byte) (ChunkResult, error) {
	var r C.aigw_chunk_result
	var ptr *C.uint8_t
	if len(b) > 0 {
		ptr = (*C.uint8_t)(unsafe.Pointer(&b[0]))
	}
	eos := C.int(0)
	if endOfStream {
		eos = 1
	}
	if rc := C.aigw_stream_chunk(s.p.ctx, s.h, ptr, C.uint32_t(len(b)), eos, (*C.uint8_t)(unsafe.Pointer(&out[0])), C.uint32_t(len(out)), &r); rc != 0 {
		return ChunkResult{}, errors.New(C.GoString(C.aigw_last_error(s.p.ctx)))
	}
	cr := ChunkResult{Status: int(r.status), Usage: r.usage, Unchanged: r.body_kind == C.AIGW_BODY_UNCHANGED}
	if !cr.Unchanged {
		cr.Body = out[:int(r.out_len)]
	}
	cr.Model = out[int(r.out_len) : int(r.out_len)+int(r.model_len)] // the response model follows the body bytes
	return cr, nil
}

func (s *Stream) Close() { C.aigw_stream_close(s.p.ctx, s.h) }

func (p *Pipeline) Close() {
	C.aigw_batcher_stop(p.batcher)
	C.aigw_destroy(p.ctx)
}
