// gen_golden.go -- emits tests/golden/chunks_go.jsonl from the REAL Go path (github.com/pbs-plus/pxar v0.19.2, reference go.mod:28;
// call site internal/pxarmount/commit.go:302-305) in the format of chunks_oracle.jsonl, so that `pytest tests/test_oracle.py -k go_golden`
// pins (or refutes) the restated oracle in one command.  Cannot be built in this image (no Go toolchain, module absent):
//	cd <pbs-plus checkout> && cp <repo>/tests/golden/gen_golden.go ./cmd/gen_golden/main.go && go run ./cmd/gen_golden -unit=kib > chunks_go.jsonl
// Two lines depend on the module's (unseen) API and are marked ADAPT.
package main

import ("crypto/sha256"; "encoding/binary"; "encoding/hex"; "encoding/json"; "flag"; "fmt"; "os"
	"github.com/pbs-plus/pxar/buzhash")

func fmix(z uint64) uint64 { z = (z ^ z>>30) * 0xBF58476D1CE4E5B9; z = (z ^ z>>27) * 0x94D049BB133111EB; return z ^ z>>31 }

// same integer recipe as oracle.c:orc_corpus_fill with dup_permille = 0, edit_mode = 0
func corpus(seed, fileID, n, blockLen uint64) []byte {
	out, bpf := make([]byte, (n+7)/8*8), (n+blockLen-1)/blockLen
	for pos := uint64(0); pos < n; pos += 8 {
		bi := pos / blockLen
		bseed := fmix(seed*0xA0761D6478BD642F + (fileID*bpf+bi)*0xE7037ED1A0B428DB + 0x1234567)
		binary.LittleEndian.PutUint64(out[pos:], fmix(bseed+((pos-bi*blockLen)/8+1)*0x9E3779B97F4A7C15))
	}
	return out[:n]
}

func main() {
	unit := flag.String("unit", "kib", "what NewConfig's argument means: kib (4096 = 4 MiB) or bytes")
	flag.Parse()
	cases := [][5]uint64{{1, 0, 1 << 20, 4096, 1 << 16}, {1, 1, 1<<20 + 12345, 4096, 1 << 16}, {2, 0, 1 << 22, 65536, 1 << 16},
		{2, 5, 300000, 1024, 1 << 12}, {3, 0, 8 << 20, 1 << 20, 1 << 20}, {1, 0, 48 << 20, 4 << 20, 4 << 20}}
	for _, c := range cases {
		arg := int(c[3]); if *unit == "kib" { arg /= 1024 }
		cfg, err := buzhash.NewConfig(arg)
		if err != nil { fmt.Fprintln(os.Stderr, "NewConfig", arg, err); continue }
		data, cuts, digs, start := corpus(c[0], c[1], c[2], c[4]), []uint64{}, []string{}, 0
		ch := buzhash.NewChunker(cfg) // ADAPT: the module's streaming chunker over cfg
		for start < len(data) {
			n := ch.Scan(data[start:]) // ADAPT: bytes up to and including the cut, 0 = no cut in this buffer
			if n == 0 { n = len(data) - start }
			d := sha256.Sum256(data[start : start+n]); start += n
			cuts, digs = append(cuts, uint64(start)), append(digs, hex.EncodeToString(d[:]))
		}
		all := sha256.Sum256(data)
		json.NewEncoder(os.Stdout).Encode(map[string]any{"gen": "pbsgpu-corpus-v1(fmix64)", "oracle": "go:" + fmt.Sprintf("%+v", cfg)[:40], "seed": c[0],
			"file_id": c[1], "len": c[2], "avg": c[3], "block_len": c[4], "data_sha256": hex.EncodeToString(all[:]), "cuts": cuts, "digests": digs})
	}
}
