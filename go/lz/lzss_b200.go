// +build cgo,raisin_b200

// Package lz — B200 build of compressor/lz.
//
// Drop-in replacement for the bodies of the three hot functions of compressor/lz/lzss.go.
// Build raisin with `-tags raisin_b200`; move the reference's own bodies of CompressAsync,
// Compress and Decompress behind `// +build !raisin_b200` (see INTEGRATION.md).  Everything
// else in lzss.go (Writer, Reader, NewWriter, NewReader, the escape helpers) stays as it is and
// keeps calling these three functions.
//
// NOTE: this image has no Go toolchain, so this file is written against the reference's
// signatures but has not been compiled here.  It only uses cgo idioms from the Go manual:
// C.GoBytes, unsafe.Pointer(&slice[0]), and no Go pointer is retained by C after the call.
package lz

/*
#cgo CFLAGS: -I${SRCDIR}/../../../include
#cgo LDFLAGS: -L${SRCDIR}/../../../raisin_b200 -lraisin_b200 -Wl,-rpath,${SRCDIR}/../../../raisin_b200
#include <stdint.h>
#include <stddef.h>
#include "raisin_b200.h"
*/
import "C"

import (
	"runtime"
	"unsafe"
)

func b200Ptr(b []byte) *C.uint8_t {
	if len(b) == 0 {
		return nil
	}
	return (*C.uint8_t)(unsafe.Pointer(&b[0]))
}

func b200Take(rc C.int, out *C.uint8_t, n C.size_t) []byte {
	if rc != 0 {
		// the reference signals every failure on this path by panicking
		panic("lzss (raisin_b200): " + C.GoString(C.rsn_strerror(rc)))
	}
	defer C.rsn_free(unsafe.Pointer(out))
	res := make([]byte, int(n))
	if n > 0 {
		// C.GoBytes takes a C.int length; copy through a big-array view so outputs > 2 GiB work
		copy(res, (*[1 << 40]byte)(unsafe.Pointer(out))[:n:n])
	}
	return res
}

// CompressAsync replaces lzss.go:109-154.  useProgressBar is cosmetic in the reference.
func CompressAsync(fileContents []byte, useProgressBar bool, maxSearchBufferLength int) []byte {
	var out *C.uint8_t
	var n C.size_t
	rc := C.rsn_lzss_compress(b200Ptr(fileContents), C.size_t(len(fileContents)),
		C.int64_t(maxSearchBufferLength), C.RSN_LZSS_ASYNC, &out, &n)
	runtime.KeepAlive(fileContents)
	return b200Take(rc, out, n)
}

// Compress replaces lzss.go:224-316 (the exported iterative variant).
func Compress(fileContents []byte, useProgressBar bool, maxSearchBufferLength int) []byte {
	var out *C.uint8_t
	var n C.size_t
	rc := C.rsn_lzss_compress(b200Ptr(fileContents), C.size_t(len(fileContents)),
		C.int64_t(maxSearchBufferLength), C.RSN_LZSS_ITER, &out, &n)
	runtime.KeepAlive(fileContents)
	return b200Take(rc, out, n)
}

// Decompress replaces lzss.go:323-364.
func Decompress(fileContents []byte, useProgressBar bool) []byte {
	var out *C.uint8_t
	var n C.size_t
	rc := C.rsn_lzss_decompress(b200Ptr(fileContents), C.size_t(len(fileContents)), &out, &n)
	runtime.KeepAlive(fileContents)
	return b200Take(rc, out, n)
}

// CompressAsyncSharded is CompressAsync for ONE large stream with the match search sharded by
// position range over ngpus GPUs of the box (BASELINE configs[4]); the bytes are those of
// CompressAsync.  Not part of the reference's package: an extra entry point for callers that hold
// very large buffers.
func CompressAsyncSharded(fileContents []byte, maxSearchBufferLength int, ngpus int) []byte {
	var out *C.uint8_t
	var n C.size_t
	rc := C.rsn_lzss_compress_sharded(b200Ptr(fileContents), C.size_t(len(fileContents)),
		C.int64_t(maxSearchBufferLength), C.RSN_LZSS_ASYNC, C.int(ngpus), &out, &n)
	runtime.KeepAlive(fileContents)
	return b200Take(rc, out, n)
}
