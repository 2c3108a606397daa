// sam2bam — SAM file -> BAM file through bam_writer.hpp (the code path `basal -o x.bam` uses). Reference names and
// lengths come from the @SQ lines. Exists so that the BAM writer can be exercised without a GPU (tests/test_bam_writer.py).
//   usage: sam2bam in.sam out.bam
#include <cstdio>
#include <fstream>
#include <iostream>
#include <sstream>

#include "bam_writer.hpp"

int main(int argc, char **argv) {
    if (argc != 3) { fprintf(stderr, "usage: sam2bam in.sam out.bam\n"); return 2; }
    std::ifstream in(argv[1]); if (!in) { fprintf(stderr, "cannot open %s\n", argv[1]); return 1; }
    std::string line, header, body; bam::Refs R;
    while (std::getline(in, line)) {
        if (!line.empty() && line[0] == '@') {
            header += line; header.push_back('\n');
            if (line.compare(0, 3, "@SQ") == 0) {
                std::string name; unsigned long len = 0; std::istringstream ss(line); std::string tok;
                while (std::getline(ss, tok, '\t')) { if (tok.compare(0, 3, "SN:") == 0) name = tok.substr(3); else if (tok.compare(0, 3, "LN:") == 0) len = strtoul(tok.c_str() + 3, nullptr, 10); }
                R.add(name, (uint32_t)len);
            }
        } else { body += line; body.push_back('\n'); }
    }
    FILE *out = fopen(argv[2], "wb"); if (!out) { fprintf(stderr, "cannot open %s\n", argv[2]); return 1; }
    std::string blk; const std::string raw = bam::header_bytes(header, R);
    bam::bgzf_append(raw.data(), raw.size(), blk);
    // several "batches", like the workers of the CLI produce them
    size_t pos = 0; const size_t chunk = 200000;
    while (pos < body.size()) {
        size_t end = std::min(body.size(), pos + chunk); while (end < body.size() && body[end - 1] != '\n') end++;
        if (!bam::text_to_blocks(body.substr(pos, end - pos), R, blk)) { fprintf(stderr, "malformed SAM record\n"); return 1; }
        pos = end;
    }
    bam::bgzf_eof(blk);
    fwrite(blk.data(), 1, blk.size(), out); fclose(out);
    return 0;
}
