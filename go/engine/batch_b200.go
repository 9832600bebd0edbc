// +build cgo,raisin_b200

// Package engine — B200 batch helper for engine.BenchmarkSuite-style loops (engine.go:208-262).
//
// The per-file stubs in ../lz and ../huffman are all a stock raisin needs; this helper is for
// callers that hold many small files and want them to go to the GPU in one call
// (rsn_batch_layers: every kernel of a stage runs once per group of files, see INTEGRATION.md).
//
// NOTE: no Go toolchain in this image — written against include/raisin_b200.h, not compiled here.
// cgo forbids passing Go memory that itself holds Go pointers, so the pointer arrays live in C
// memory and the file bytes are copied into library-owned pinned buffers (rsn_host_alloc) first.
package engine

/*
#cgo CFLAGS: -I${SRCDIR}/../../../include
#cgo LDFLAGS: -L${SRCDIR}/../../../raisin_b200 -lraisin_b200 -Wl,-rpath,${SRCDIR}/../../../raisin_b200
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include "raisin_b200.h"
*/
import "C"

import (
	"errors"
	"strings"
	"unsafe"
)

// BatchLayers runs every file through the layer list ("lzss", "huffman") as engine.compress /
// engine.decompress would (engine.go:443-479).  errs[i] != nil where the reference would have
// panicked on file i (AsyncBenchmarkFile turns that into Result.Failed, engine.go:344-350).
func BatchLayers(algorithms []string, compress bool, files [][]byte) (outs [][]byte, errs []error) {
	n := len(files)
	outs = make([][]byte, n)
	errs = make([]error, n)
	if n == 0 {
		return
	}
	calgos := C.CString(strings.Join(algorithms, ","))
	defer C.free(unsafe.Pointer(calgos))
	psz := C.size_t(unsafe.Sizeof(uintptr(0)))
	ins := (*[1 << 28]*C.uint8_t)(C.calloc(C.size_t(n), psz))[:n:n]
	res := (*[1 << 28]*C.uint8_t)(C.calloc(C.size_t(n), psz))[:n:n]
	ns := (*[1 << 28]C.size_t)(C.calloc(C.size_t(n), C.size_t(unsafe.Sizeof(C.size_t(0)))))[:n:n]
	resN := (*[1 << 28]C.size_t)(C.calloc(C.size_t(n), C.size_t(unsafe.Sizeof(C.size_t(0)))))[:n:n]
	rcs := (*[1 << 28]C.int)(C.calloc(C.size_t(n), C.size_t(unsafe.Sizeof(C.int(0)))))[:n:n]
	defer func() {
		for i := range ins {
			if ins[i] != nil {
				C.rsn_host_free(unsafe.Pointer(ins[i]))
			}
		}
		C.free(unsafe.Pointer(&ins[0]))
		C.free(unsafe.Pointer(&res[0]))
		C.free(unsafe.Pointer(&ns[0]))
		C.free(unsafe.Pointer(&resN[0]))
		C.free(unsafe.Pointer(&rcs[0]))
	}()
	for i, f := range files {
		ns[i] = C.size_t(len(f))
		if len(f) > 0 {
			p := C.rsn_host_alloc(C.size_t(len(f)))
			if p == nil {
				errs[i] = errors.New("raisin_b200: out of pinned host memory")
				ns[i] = 0
				continue
			}
			copy((*[1 << 40]byte)(p)[:len(f):len(f)], f)
			ins[i] = (*C.uint8_t)(p)
		}
	}
	flag := C.int(0)
	if compress {
		flag = 1
	}
	const notReached = -1000 // no code of the library: the call failed before it came to this file
	for i := range rcs {
		rcs[i] = notReached
	}
	rc := C.rsn_batch_layers(calgos, flag, C.size_t(n), (**C.uint8_t)(unsafe.Pointer(&ins[0])), &ns[0],
		(**C.uint8_t)(unsafe.Pointer(&res[0])), &resN[0], &rcs[0], 0, 0)
	for i := 0; i < n; i++ {
		if rcs[i] == notReached { // unknown layer name, no device, ...: every file fails with the call's code
			errs[i] = errors.New("raisin_b200: " + C.GoString(C.rsn_strerror(rc)))
			continue
		}
		if rcs[i] != 0 {
			if errs[i] == nil {
				errs[i] = errors.New("raisin_b200: " + C.GoString(C.rsn_strerror(rcs[i])))
			}
			continue
		}
		outs[i] = make([]byte, int(resN[i]))
		if resN[i] > 0 {
			copy(outs[i], (*[1 << 40]byte)(unsafe.Pointer(res[i]))[:resN[i]:resN[i]])
		}
		C.rsn_free(unsafe.Pointer(res[i]))
	}
	return
}

// BenchmarkFileB200 is engine.BenchmarkFile (engine.go:357-441) with the histograms, the entropies
// and the lossless comparison computed on the device in the same call as the layers
// (rsn_benchmark_file).  The fields are those of engine.Result.
func BenchmarkFileB200(algorithms []string, fileContents []byte) (timeTaken float64, ratio float32, actualEntropy float32,
	entropy float64, lossless bool, failed bool) {
	calgos := C.CString(strings.Join(algorithms, ","))
	defer C.free(unsafe.Pointer(calgos))
	var r C.rsn_bench_result
	var p *C.uint8_t
	if len(fileContents) > 0 {
		p = (*C.uint8_t)(unsafe.Pointer(&fileContents[0]))
	}
	if rc := C.rsn_benchmark_file(calgos, p, C.size_t(len(fileContents)), &r); rc != 0 {
		panic("raisin_b200: " + C.GoString(C.rsn_strerror(rc)))
	}
	return float64(r.seconds), float32(r.ratio), float32(r.actual_entropy), float64(r.entropy), r.lossless != 0, r.failed != 0
}
