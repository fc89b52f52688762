// Read files as a byte stream for the record reader (A0, parse_reads GC.cpp:1220-1287).  seqan3 opens read files through
// its transparent decompression layer (seqan3/io/detail/misc_input.hpp: magic-header detection, then zlib's inflate on one
// thread, or its BGZF stream); here plain files are read with parallel preads and gzip files -- single-member, multi-member
// or BGZF alike -- are inflated by all host threads at once:
//   * the compressed stream is cut into ranges of 1 MiB; for every range but the first a worker looks for the first deflate
//     block that starts with a valid dynamic-Huffman header (bit-granular search without a branch per position, RFC 1951
//     3.2.7) and another decodes from there with an UNKNOWN 32 KiB history: output symbols are 16 bits wide, a
//     back-reference that reaches before the chunk's start yields a marker naming the history position it wants (the
//     two-pass scheme of pugz / rapidgzip);
//   * a chunk's decoder stops at the block boundary where a later range's decoder started (a candidate it runs past was
//     a false positive and is dropped: the predecessor simply keeps decoding through it);
//   * finders, decoders and the consumer's work are tasks of one scheduler taken in stream order (no barriers: a worker
//     that finishes a chunk takes the next one); a sequencer thread follows the chain of chunks, propagates the
//     histories (32 KiB each) and decodes a chunk itself where the chain breaks (stored / fixed-Huffman blocks the finder
//     does not look for, a block longer than a decoder's view);
//   * markers are replaced by the workers in slices of 1 MiB straight into the caller's buffer, and the CRC-32 / ISIZE
//     trailer of every gzip member is verified from the slices' CRCs (carry-less multiplication, crc32_combine).
// Any structural error or CRC mismatch is reported (GNB_ERR_IO), as zlib would.
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <string>

namespace gnb
{

class ThreadPool;

class ByteSource
{
  public:
    virtual ~ByteSource() = default;
    // next bytes of the (decompressed) stream: > 0 bytes written to dst, 0 at the end, < 0 = gnb_status (message in error())
    virtual int64_t read(char *dst, size_t cap) = 0;
    // plain files only: the stream position can be set (sliced ingest re-reads from the consumed position)
    virtual bool    seekable() const { return false; }
    virtual int64_t read_at(char *dst, size_t cap, uint64_t offset) { (void)dst, (void)cap, (void)offset; return -1; }
    virtual uint64_t size() const { return 0; }
    virtual bool    is_gzip() const { return false; }
    const std::string &error() const { return err_; }

  protected:
    std::string err_;
};

// bzip2 files ("BZh"): the blocks decoded by `threads` host threads through libbz2 (bz2stream.cpp)
std::unique_ptr<ByteSource> open_bz2_source(int fd, uint64_t size, int threads);

// plain, gzip (1f 8b) or bzip2 ("BZh") by magic number, like seqan3's make_secondary_istream
// threads: workers of the source (0 = as many as the host suggests, divided by `share` = files read at the same time)
// (files named .embl / .genbank / .gb / .gbk / .sam, compressed or not, come out rewritten as two-line FASTA: seqformats.cpp)
std::unique_ptr<ByteSource> open_byte_source(const std::string &path, int threads, std::string &err, int share = 1);

// seqan3::sequence_file_input picks the format from the file name, compression suffix stripped first (`file_extensions` of
// format_fasta.hpp / format_fastq.hpp / format_embl.hpp / format_genbank.hpp / format_sam.hpp).  Unknown: seqan3 throws
// unhandled_extension_error, which ganon-classify does not catch.
enum
{
    kFormatUnknown = 0,
    kFormatFasta   = 1,
    kFormatFastq   = 2,
    kFormatEmbl    = 3,
    kFormatGenbank = 4,
    kFormatSam     = 5
};
int format_of_extension(std::string name);
// EMBL / GenBank / SAM by file name: the same stream rewritten record by record as two-line FASTA; anything else: src itself
std::unique_ptr<ByteSource> wrap_sequence_format(std::unique_ptr<ByteSource> src, const std::string &path);

} // namespace gnb
