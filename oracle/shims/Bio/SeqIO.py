"""SeqIO.parse(handle, "fasta") -> records with a `.seq` that str() turns into the sequence (search_organisms.py:946-948)."""


class _Record:
    def __init__(self, name, seq):
        self.id = self.name = self.description = name
        self.seq = seq


def parse(handle, fmt):
    if fmt != "fasta":
        raise ValueError("this stand-in reads FASTA only")
    name, chunks = None, []
    for line in handle:
        line = line.strip()
        if line.startswith(">"):
            if name is not None:
                yield _Record(name, "".join(chunks))
            name, chunks = line[1:], []
        elif line:
            chunks.append(line)
    if name is not None:
        yield _Record(name, "".join(chunks))
