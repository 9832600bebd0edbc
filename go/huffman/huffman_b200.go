// +build cgo,raisin_b200

// Package huffman — B200 build of compressor/huffman.
//
// Replaces the bodies of Compress (huffman.go:299-325) and Decompress (huffman.go:327-330).
// Writer/Reader/NewWriter/NewReader (huffman.go:368-422) stay as they are.  The package-level
// builders `estring` and `answer` (huffman.go:56,129) are no longer touched, which also removes
// the data race between BenchmarkSuite goroutines (engine/engine.go:243).
//
// NOTE: not compiled in this image (no Go toolchain); see INTEGRATION.md.
package huffman

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

// StrictLimits reproduces the reference's 900000-bit recursion guard (huffman.go:132-134)
// when true.  Default false: inputs the reference cannot decode (> ~110 KiB of text) work.
var StrictLimits = false

func b200Ptr(b []byte) *C.uint8_t {
	if len(b) == 0 {
		return nil
	}
	return (*C.uint8_t)(unsafe.Pointer(&b[0]))
}

func b200Take(rc C.int, out *C.uint8_t, n C.size_t) []byte {
	if rc != 0 {
		panic("huffman (raisin_b200): " + C.GoString(C.rsn_strerror(rc)))
	}
	defer C.rsn_free(unsafe.Pointer(out))
	res := make([]byte, int(n))
	if n > 0 {
		// C.GoBytes takes a C.int length; copy through a big-array view so outputs > 2 GiB work
		copy(res, (*[1 << 40]byte)(unsafe.Pointer(out))[:n:n])
	}
	return res
}

// Compress replaces huffman.go:299-325.
func Compress(fileContents []byte) []byte {
	var out *C.uint8_t
	var n C.size_t
	rc := C.rsn_huff_compress(b200Ptr(fileContents), C.size_t(len(fileContents)), &out, &n)
	runtime.KeepAlive(fileContents)
	return b200Take(rc, out, n)
}

// Decompress replaces huffman.go:327-330.
func Decompress(fileContents []byte) []byte {
	var out *C.uint8_t
	var n C.size_t
	strict := C.int(0)
	if StrictLimits {
		strict = 1
	}
	rc := C.rsn_huff_decompress(b200Ptr(fileContents), C.size_t(len(fileContents)), strict, &out, &n)
	runtime.KeepAlive(fileContents)
	return b200Take(rc, out, n)
}
