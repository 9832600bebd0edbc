// Dumps what the STOCK reference produces for the inputs of tests/golden/inputs.json.gz, in the format
// of tests/golden/vectors.json, so the oracle-derived vectors can be checked against the real Go
// code ("parity unpinned" in DESIGN.md: this image has no Go toolchain, so this has never been run
// here).  In a raisin checkout, with this file copied to compressor/golden_dump_test.go:
//
//	RAISIN_B200_INPUTS=/path/to/tests/golden/inputs.json.gz \
//	RAISIN_B200_OUT=/path/to/reference_outputs.json go test ./compressor -run TestDumpGolden
//	python tests/tools/compare_reference_outputs.py /path/to/reference_outputs.json
//
// Huffman headers come out in Go map order, which is random: the comparison script compares the
// payload bytes, the header as a set of records and the total length, not the header bytes.
package compressor_test

import (
	"bytes"
	"compress/gzip"
	"crypto/sha256"
	"encoding/hex"
	"encoding/json"
	"fmt"
	"io/ioutil"
	"os"
	"testing"

	"github.com/go-compression/raisin/compressor/huffman"
	"github.com/go-compression/raisin/compressor/lz"
)

type rec map[string]interface{}

func record(b []byte) rec {
	sum := sha256.Sum256(b)
	r := rec{"len": len(b), "sha256": hex.EncodeToString(sum[:])}
	if len(b) <= 96 {
		r["hex"] = hex.EncodeToString(b)
	}
	return r
}

// the reference signals every failure on this path by panicking
func attempt(f func() []byte) (r rec) {
	defer func() {
		if e := recover(); e != nil {
			r = rec{"error": fmt.Sprint(e)}
		}
	}()
	return record(f())
}

func TestDumpGolden(t *testing.T) {
	raw, err := ioutil.ReadFile(os.Getenv("RAISIN_B200_INPUTS"))
	if err != nil {
		t.Skip("RAISIN_B200_INPUTS not set")
	}
	if zr, err := gzip.NewReader(bytes.NewReader(raw)); err == nil { // inputs.json.gz
		if raw, err = ioutil.ReadAll(zr); err != nil {
			t.Fatal(err)
		}
	}
	var in map[string]map[string]string
	if err := json.Unmarshal(raw, &in); err != nil {
		t.Fatal(err)
	}
	out := map[string]map[string]rec{"lzss": {}, "huffman": {}}
	for name, hx := range in["lzss"] {
		data, _ := hex.DecodeString(hx)
		e := rec{"input": record(data)}
		e["async_w4096"] = attempt(func() []byte { return lz.CompressAsync(data, false, 4096) })
		e["async_w1024"] = attempt(func() []byte { return lz.CompressAsync(data, false, 1024) })
		e["iter_w4096"] = attempt(func() []byte { return lz.Compress(data, false, 4096) })
		e["decompress_async"] = attempt(func() []byte { return lz.Decompress(lz.CompressAsync(data, false, 4096), false) })
		e["decompress_iter"] = attempt(func() []byte { return lz.Decompress(lz.Compress(data, false, 4096), false) })
		e["layered_full"] = attempt(func() []byte { return huffman.Compress(lz.CompressAsync(data, false, 4096)) })
		out["lzss"][name] = e
	}
	for name, hx := range in["huffman"] {
		data, _ := hex.DecodeString(hx)
		e := rec{"input": record(data)}
		e["compressed_full"] = attempt(func() []byte { return huffman.Compress(data) })
		e["decompress"] = attempt(func() []byte { return huffman.Decompress(huffman.Compress(data)) })
		out["huffman"][name] = e
	}
	// full bytes of the Huffman outputs are needed to split header and payload
	full := map[string]string{}
	for name, hx := range in["huffman"] {
		data, _ := hex.DecodeString(hx)
		func() {
			defer func() { recover() }()
			full["huffman/"+name] = hex.EncodeToString(huffman.Compress(data))
		}()
	}
	for name, hx := range in["lzss"] {
		data, _ := hex.DecodeString(hx)
		func() {
			defer func() { recover() }()
			full["layered/"+name] = hex.EncodeToString(huffman.Compress(lz.CompressAsync(data, false, 4096)))
		}()
	}
	blob, _ := json.MarshalIndent(map[string]interface{}{"vectors": out, "huffman_full_hex": full}, "", " ")
	if err := ioutil.WriteFile(os.Getenv("RAISIN_B200_OUT"), blob, 0o644); err != nil {
		t.Fatal(err)
	}
}
