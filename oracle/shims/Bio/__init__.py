"""Stand-in for Biopython: only SeqIO.parse(handle, "fasta") is provided (see ../README.md)."""
