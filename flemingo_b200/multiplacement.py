"""Drop-in for the reference's `_multiplacement` extension module (src/_interface.c:187-208).

    import flemingo_b200.multiplacement as _multiplacement
    _multiplacement.calculate(sequence, rec_types, rec_matrices, rec_lengths, con_matrices,
                              rec_scores, con_scores, con_lengths, max_length, bin_freqs, bin_edges, num_bins)

Same positional / keyword names (kwlist, _interface.c:49-52), same in-place outputs, returns None.  The work is
done by fl_calculate (one placement on the GPU).  Differences, all on error paths:
  * an organism longer than the sequence raises ValueError (the reference sets RuntimeError while returning
    None, _interface.c:162-172 -- a latent SystemError);
  * a NaN null-model bin raises instead of calling exit(1) (recognizer.c:258-262);
  * PRECOMPUTE=false connector arrays are rejected (legacy float maths, connector.c:70-72).
"""
from __future__ import annotations

from .engine import get_engine


def calculate(sequence, rec_types, rec_matrices, rec_lengths, con_matrices, rec_scores, con_scores, con_lengths,
              max_length, bin_freqs, bin_edges, num_bins) -> None:
    get_engine().calculate(sequence, rec_types, rec_matrices, rec_lengths, con_matrices, rec_scores, con_scores,
                           con_lengths, max_length, bin_freqs, bin_edges, num_bins)
