//go:build gpu && cgo

// Package gpuchunk is the cgo shim a pbs-plus maintainer adds to call libpbsgpu.so
// (include/pbsgpu.h) from the Go host code.  It is the reference-side binding for the
// ONE hot path this repository accelerates: buzhash boundary scan + per-chunk SHA-256 +
// known-digest probe, i.e. what happens inside
//
//	writer.WriteEntryReader(entry, tee, size)          internal/pxarmount/commit.go:720, :858
//
// configured by buzhash.NewConfig(4096) (commit.go:302-305) and seeded from the previous
// snapshot's index (commit.go:286-294, :324-329).
//
// NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Go toolchain (SURVEY.md
// section 0).  The same C symbols are exercised by the Python ctypes tests (tests/) and the
// C++ driver (tests/cxx/).  pxar-mount is built CGO_ENABLED=0 today (.goreleaser.yaml:56);
// this file only builds with `-tags gpu` and CGO_ENABLED=1.
package gpuchunk

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../pbs_plus_b200 -lpbsgpu -Wl,-rpath,${SRCDIR}/../../pbs_plus_b200
#include <stdlib.h>
#include "pbsgpu.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"io"
	"runtime"
	"unsafe"
)

// Config mirrors buzhash.Config (opaque in the reference; built by NewConfig).
type Config struct{ c C.pbsgpu_cfg }

// NewConfig mirrors buzhash.NewConfig(avgKiB) (commit.go:303 passes 4096 = 4 MiB).
func NewConfig(avgKiB int) (Config, error) {
	var cfg Config
	if rc := C.pbsgpu_config_kib(C.uint32_t(avgKiB), nil, &cfg.c); rc != 0 {
		return cfg, fmt.Errorf("buzhash: invalid average chunk size %d KiB (rc %d)", avgKiB, int(rc))
	}
	return cfg, nil
}

// Engine is one GPU context.  Safe to use from any goroutine: the C side binds the device
// per call and serialises calls per context (no thread-local CUDA state).
type Engine struct{ ctx *C.pbsgpu_ctx }

func Open(device int) (*Engine, error) {
	var ctx *C.pbsgpu_ctx
	if rc := C.pbsgpu_open(C.int(device), &ctx); rc != 0 {
		return nil, fmt.Errorf("pbsgpu_open(%d): rc %d (no CUDA device; there is no CPU fallback)", device, int(rc))
	}
	e := &Engine{ctx}
	runtime.SetFinalizer(e, func(e *Engine) { e.Close() })
	return e, nil
}

func (e *Engine) Close() {
	if e.ctx != nil {
		C.pbsgpu_close(e.ctx)
		e.ctx = nil
	}
}

func (e *Engine) err(rc C.int) error {
	if rc == 0 {
		return nil
	}
	return fmt.Errorf("pbsgpu: %s (rc %d)", C.GoString(C.pbsgpu_strerror(e.ctx)), int(rc))
}

// KnownSet mirrors the session's known-chunk bookkeeping (PreviousBackupRef, commit.go:286-294).
type KnownSet struct {
	e *Engine
	s *C.pbsgpu_set
}

func (e *Engine) NewKnownSet(capacityHint uint64) (*KnownSet, error) {
	var s *C.pbsgpu_set
	if err := e.err(C.pbsgpu_set_create(e.ctx, C.uint64_t(capacityHint), &s)); err != nil {
		return nil, err
	}
	return &KnownSet{e, s}, nil
}

// SeedFromDidx feeds the bytes of the previous .ppxar.didx (origPayloadIdx, commit.go:324-328).
func (k *KnownSet) SeedFromDidx(didx []byte) (uint64, error) {
	if len(didx) == 0 {
		return 0, nil
	}
	var n C.uint64_t
	rc := C.pbsgpu_set_seed_didx(k.s, (*C.uint8_t)(unsafe.Pointer(&didx[0])), C.uint64_t(len(didx)), &n)
	return uint64(n), k.e.err(rc)
}

func (k *KnownSet) Close() { C.pbsgpu_set_destroy(k.s) }

// Chunk is one dynamic-index entry: (end offset, digest) + whether the digest was known.
type Chunk struct {
	Stream uint32
	Known  bool
	End    uint64
	Digest [32]byte
}

// ErrStagingFull is returned by WriteEntryReader when the pinned staging buffer cannot take the entry:
// Flush the batch and write the entry again.
var ErrStagingFull = errors.New("pbsgpu: staging full, call Flush first")

// Batch accumulates whole files in C-owned PINNED staging (Go pointers are never retained by C)
// and pushes them through the GPU in one call -- the batched form of the per-file loop at
// commit.go:604-625 / :697-731.
type Batch struct {
	e    *Engine
	cfg  Config
	buf  unsafe.Pointer
	cap  uint64
	fill uint64
	off  []C.uint64_t
	ln   []C.uint64_t
}

func (e *Engine) NewBatch(cfg Config, stagingBytes uint64) (*Batch, error) {
	p := C.pbsgpu_host_alloc(e.ctx, C.uint64_t(stagingBytes))
	if p == nil {
		return nil, errors.New("pbsgpu: pinned staging allocation failed")
	}
	return &Batch{e: e, cfg: cfg, buf: p, cap: stagingBytes}, nil
}

// WriteEntryReader mirrors transfer.ArchiveWriter.WriteEntryReader(entry, reader, size): it pulls
// exactly size bytes from r (io.ReadFull semantics) into the staging buffer.
func (b *Batch) WriteEntryReader(r io.Reader, size uint64) error {
	start := (b.fill + 255) &^ 255
	if start+size > b.cap {
		return ErrStagingFull
	}
	dst := unsafe.Slice((*byte)(unsafe.Add(b.buf, start)), size)
	if _, err := io.ReadFull(r, dst); err != nil {
		return fmt.Errorf("read payload: %w", err)
	}
	b.off = append(b.off, C.uint64_t(start))
	b.ln = append(b.ln, C.uint64_t(size))
	b.fill = start + size
	return nil
}

// Flush runs scan -> cut -> SHA-256 -> probe for every queued file and returns the chunks in
// (file, offset) order; Known chunks need no upload ("Only new chunks are uploaded").
func (b *Batch) Flush(known *KnownSet) ([]Chunk, error) {
	res, _, err := b.flush(known, false)
	return res, err
}

// FlushWithHashes is Flush plus the XXH3-64 (seed 0) of every queued file, computed on the GPU from the
// same staged bytes: the value emitBackedFile gets from `h := xxh3.New(); io.TeeReader(f, h); h.Sum64()`
// (commit.go:717-725) and stores in ow.backedHashes -- so the host no longer hashes while it reads.
func (b *Batch) FlushWithHashes(known *KnownSet) ([]Chunk, []uint64, error) {
	return b.flush(known, true)
}

func (b *Batch) flush(known *KnownSet, withHashes bool) ([]Chunk, []uint64, error) {
	n := len(b.off)
	if n == 0 {
		return nil, nil, nil
	}
	capChunks := uint64(n)
	for _, l := range b.ln {
		capChunks += uint64(l) / uint64(b.cfg.c.min)
	}
	out := make([]C.pbsgpu_chunk, capChunks+1)
	var nOut C.uint64_t
	var set *C.pbsgpu_set
	if known != nil {
		set = known.s
	}
	var hashes []C.uint64_t
	var hp *C.uint64_t
	if withHashes {
		hashes = make([]C.uint64_t, n)
		hp = &hashes[0]
	}
	rc := C.pbsgpu_chunk_digest_batch_xxh3(b.e.ctx, &b.cfg.c, b.buf, &b.off[0], &b.ln[0], C.uint32_t(n), set,
		&out[0], C.uint64_t(len(out)), &nOut, hp)
	if err := b.e.err(rc); err != nil {
		return nil, nil, err
	}
	res := make([]Chunk, int(nOut))
	for i := range res {
		res[i].Stream = uint32(out[i].stream)
		res[i].Known = out[i].flags&C.PBSGPU_CHUNK_KNOWN != 0
		res[i].End = uint64(out[i].end_off)
		copy(res[i].Digest[:], C.GoBytes(unsafe.Pointer(&out[i].digest[0]), 32))
	}
	var hs []uint64
	if withHashes {
		hs = make([]uint64, n)
		for i := range hashes {
			hs[i] = uint64(hashes[i])
		}
	}
	b.off, b.ln, b.fill = b.off[:0], b.ln[:0], 0
	return res, hs, nil
}

// FileHashes returns the XXH3-64 of every queued file WITHOUT chunking: the verify pass of the commit
// (verifyBackedFileHashes, commit.go:957-976) re-reads the backed files into the staging buffer with
// WriteEntryReader and compares these values with ow.backedHashes.  The queue is cleared.
func (b *Batch) FileHashes() ([]uint64, error) {
	n := len(b.off)
	if n == 0 {
		return nil, nil
	}
	hashes := make([]C.uint64_t, n)
	rc := C.pbsgpu_xxh3_batch(b.e.ctx, b.buf, &b.off[0], &b.ln[0], C.uint32_t(n), &hashes[0])
	if err := b.e.err(rc); err != nil {
		return nil, err
	}
	res := make([]uint64, n)
	for i := range hashes {
		res[i] = uint64(hashes[i])
	}
	b.off, b.ln, b.fill = b.off[:0], b.ln[:0], 0
	return res, nil
}

func (b *Batch) Close() { C.pbsgpu_host_free(b.e.ctx, b.buf) }

// BuildDidx renders the dynamic-index image (<name>.ppxar.didx, commit.go:321-322) for chunks in
// (stream, offset) order; offsets are cumulative over the archive stream, the checksum is computed on the GPU.
func (e *Engine) BuildDidx(chunks []Chunk, uuid [16]byte, ctime int64) ([]byte, error) {
	n := len(chunks)
	rec := make([]C.pbsgpu_chunk, n+1)
	for i, c := range chunks {
		rec[i].stream = C.uint32_t(c.Stream)
		rec[i].end_off = C.uint64_t(c.End)
		for k := 0; k < 32; k++ {
			rec[i].digest[k] = C.uint8_t(c.Digest[k])
		}
	}
	out := make([]byte, uint64(C.pbsgpu_didx_size(C.uint64_t(n))))
	rc := C.pbsgpu_didx_build(e.ctx, &rec[0], C.uint64_t(n), (*C.uint8_t)(unsafe.Pointer(&uuid[0])), C.int64_t(ctime),
		(*C.uint8_t)(unsafe.Pointer(&out[0])), C.uint64_t(len(out)))
	return out, e.err(rc)
}

// BlobCRC32 returns the DataBlob payload checksums of ranges [off[i], off[i]+len[i]) of the batch's pinned
// staging buffer (only NEW chunks need a blob; POST /dynamic_chunk).
func (b *Batch) BlobCRC32(off, length []uint64) ([]uint32, error) {
	n := len(off)
	if n == 0 {
		return nil, nil
	}
	o := make([]C.uint64_t, n)
	l := make([]C.uint64_t, n)
	for i := range off {
		o[i], l[i] = C.uint64_t(off[i]), C.uint64_t(length[i])
	}
	crc := make([]C.uint32_t, n)
	rc := C.pbsgpu_crc32_batch(b.e.ctx, b.buf, &o[0], &l[0], C.uint32_t(n), &crc[0])
	if err := b.e.err(rc); err != nil {
		return nil, err
	}
	res := make([]uint32, n)
	for i := range crc {
		res[i] = uint32(crc[i])
	}
	return res, nil
}
