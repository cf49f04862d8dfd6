"""FASTA / FASTQ files at the corners of kseq's record rules (kseq.h:168-207), shared by the tests that pin the oracle's
parser on the reference binary and the drivers' reader on the oracle."""
EDGE = {
    "empty.fa": b"",
    "no_newline.fa": b">a\nACGT\n>b x y\nAC\nGT",
    "header_only.fa": b">a\n>b\nAC\n>c",
    "blank_lines.fa": b"\n\n>a\n\nAC\n\nGT\n\n>b\n\n",
    "crlf.fq": b"@a\r\nACGT\r\n+\r\nIIII\r\n@b\r\nAC\r\n+\r\nII\r\n",
    "multi.fq": b"@a\nACGT\nACG\n+a\n@III\n>II\n@b\nA\n+\nI\n",
    "truncated.fq": b"@a\nACGT\n+\nIIII\n@b\nACGT\n+\nII",
    "no_qual.fq": b"@a\nACGT\n+\nIIII\n@b\nACGT\n+",
    "junk_first.fa": b"junk line\nmore junk\n>a\nACGT\n",
    "mixed.fx": b">a\nACGT\n@b\nACGT\n+\nIIII\n>c\nAC\n",
}
