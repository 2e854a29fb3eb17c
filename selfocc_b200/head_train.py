"""Training-form forward of NeuSHead (neus_head.py:473-713).  Implemented in a later milestone."""


def forward_train(head, representation, metas=None, **kwargs):
    raise NotImplementedError('NeuSHead.forward (training form with per-sample outputs) is not built yet; '
                              'use prepare()+render() or forward_occ()')
