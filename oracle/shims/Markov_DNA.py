"""Stand-in for Markov_DNA.MCM (search_organisms.py:573-582): an order-k Markov chain trained on one sequence.
Only used when NEGATIVE_FILENAME is null; the test configuration names a negative FASTA file instead."""
import random


class MCM:
    def __init__(self, order):
        self.k = order
        self.counts = {}
        self.seq = ""

    def train(self, seq):
        self.seq = seq
        self.counts = {}
        for i in range(len(seq) - self.k):
            self.counts.setdefault(seq[i:i + self.k], []).append(seq[i + self.k])

    def generate(self, size, N=1):
        out = []
        for _ in range(N):
            start = random.randint(0, max(0, len(self.seq) - self.k)) if self.seq else 0
            s = list(self.seq[start:start + self.k])
            while len(s) < size:
                nxt = self.counts.get("".join(s[len(s) - self.k:]) if self.k else "", None)
                s.append(random.choice(nxt) if nxt else random.choice(self.seq or "ACGT"))
            out.append("".join(s[:size]))
        return out
