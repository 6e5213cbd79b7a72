"""CPU-only behaviour tests of the optimizer front-end's bookkeeping and of the Updater's dispatch
(python/mxnet/optimizer/optimizer.py:412-509, 2071-2176 semantics). No operator runs: the Updater is
driven with a recording optimizer and stand-in arrays."""
import pickle

import numpy as np
import pytest
from compat import mxnet_optimizer as mxopt   # the reference's optimizer front-end, mirrored (test infrastructure)


@pytest.fixture(scope="module")
def mx():
    import anand_mxnet_b200 as mx
    return mx


class _Ctx(object):
    def __init__(self, dev):
        self.device_id = dev


class _Arr(object):
    """stand-in for an NDArray: what Updater.__call__ looks at"""
    def __init__(self, name, dtype=np.float32, dev=0):
        self.name, self.dtype, self.context = name, np.dtype(dtype).type, _Ctx(dev)

    def __repr__(self):
        return self.name


def test_multiplier_precedence_and_counts(mx):
    class P(object):
        def __init__(self, lr_mult, wd_mult):
            self.lr_mult, self.wd_mult = lr_mult, wd_mult
    opt = mxopt.SGD(learning_rate=0.5, wd=0.1, begin_num_update=10,
                           param_idx2name={0: 'a_weight', 1: 'a_bias', 2: 'b_gamma', 3: 'c_beta'},
                           param_dict={3: P(7.0, 3.0)})
    # wd: names not ending in _weight / _gamma get multiplier 0; param_dict wins over names
    assert opt._get_wds([0, 1, 2, 3]) == [0.1, 0.0, 0.1, 0.1 * 3.0]
    opt.set_lr_mult({'a_weight': 2.0, 1: 4.0})          # by name and by index
    assert opt._get_lrs([0, 1, 2, 3]) == [1.0, 2.0, 0.5, 3.5]
    assert opt._get_lr(1) == 2.0 and opt._get_wd(3) == pytest.approx(0.3)
    opt.set_wd_mult({'a_bias': 0.5})
    assert opt._get_wds([1]) == [0.05]
    # update counts start at begin_num_update, are kept per device, num_update is the max seen
    opt._update_count([0, 1])
    opt._update_count(0)
    assert opt._index_update_count == {0: 12, 1: 11} and opt.num_update == 12
    opt._set_current_context(1)
    opt._update_count(0)
    assert opt._index_update_count == {0: 11} and opt.num_update == 12
    opt._set_current_context(0)
    assert opt._index_update_count == {0: 12, 1: 11}


def test_scheduler_and_learning_rate_rules(mx, capsys):
    class Sched(object):
        base_lr = 0.3

        def __call__(self, n):
            return 1.0 / (1 + n)
    s = Sched()
    opt = mxopt.SGD(learning_rate=0.2, lr_scheduler=s)
    assert s.base_lr == 0.2                       # overwritten, with a warning printed
    assert 'overwritten' in capsys.readouterr().out
    assert opt.learning_rate == 1.0 and opt._get_lrs([5]) == [1.0]
    opt.num_update = 3
    assert opt.learning_rate == 0.25
    with pytest.raises(UserWarning):
        opt.set_learning_rate(0.1)
    plain = mxopt.SGD()
    assert plain.lr == 0.01                       # default when neither lr nor scheduler is given
    plain.set_learning_rate(0.7)
    assert plain.learning_rate == 0.7
    st = pickle.loads(pickle.dumps(plain))
    assert st.lr == 0.7 and st.param_dict == {}


def test_registry(mx):
    assert mxopt.create('SGD').__class__ is mxopt.SGD
    with pytest.raises(ValueError):
        mxopt.create('no_such_optimizer')
    with pytest.warns(UserWarning):
        @mxopt.register
        class sgd(mxopt.SGD):  # noqa: N801 - same registry name on purpose
            pass
    mxopt.register(mxopt.SGD)       # restore


def _recording_optimizer(mx, aggregate_num):
    class Rec(mxopt.Optimizer):
        def __init__(self):
            super(Rec, self).__init__(learning_rate=0.1)
            self.aggregate_num = aggregate_num
            self.calls, self.created = [], []

        def create_state_multi_precision(self, index, weight):
            self.created.append(index)
            return 'state-%s' % index

        def update_multi_precision(self, index, weight, grad, state):
            self.calls.append((index, weight, grad, state))
    return Rec()


def test_updater_aggregates_by_dtype_in_chunks(mx):
    opt = _recording_optimizer(mx, 2)
    upd = mxopt.get_updater(opt)
    ws = [_Arr('w0'), _Arr('w1', np.float16), _Arr('w2'), _Arr('w3'), _Arr('w4', np.float16)]
    gs = [_Arr('g%d' % i) for i in range(5)]
    upd([0, 1, 2, 3, 4], gs, ws)
    assert opt.created == [0, 1, 2, 3, 4]
    # float32 group {0,2,3} in chunks of 2, then the float16 group {1,4}; order inside a group kept
    got = [(c[0], [w.name for w in c[1]], [g.name for g in c[2]], c[3]) for c in opt.calls]
    assert got == [([0, 2], ['w0', 'w2'], ['g0', 'g2'], ['state-0', 'state-2']),
                   ([3], ['w3'], ['g3'], ['state-3']),
                   ([1, 4], ['w1', 'w4'], ['g1', 'g4'], ['state-1', 'state-4'])]
    # states are created once
    upd([0, 2], [gs[0], gs[2]], [ws[0], ws[2]])
    assert opt.created == [0, 1, 2, 3, 4] and opt.calls[-1][0] == [0, 2]
    # bytes keys are decoded
    upd([b'k'], [gs[0]], [ws[0]])
    assert opt.created[-1] == 'k'


def test_updater_without_aggregation_and_device_counters(mx):
    opt = _recording_optimizer(mx, 0)
    upd = mxopt.get_updater(opt)
    upd(7, _Arr('g'), _Arr('w', dev=3))
    assert opt.calls == [(7, opt.calls[0][1], opt.calls[0][2], 'state-7')]
    assert opt.calls[0][1].name == 'w' and 3 in opt._all_index_update_counts
    upd([1, 2], [_Arr('ga'), _Arr('gb')], [_Arr('wa'), _Arr('wb')])
    assert [c[0] for c in opt.calls] == [7, 1, 2]           # one call per key


def test_updater_states_round_trip(mx):
    opt = _recording_optimizer(mx, 0)
    upd = mxopt.get_updater(opt)
    upd.states = {0: ('m', 'v'), 'k': None}
    blob = upd.get_states()
    other = mxopt.get_updater(mxopt.SGD(learning_rate=0.3))
    other.set_states(blob)
    assert other.states == {0: ('m', 'v'), 'k': None} and other.optimizer.lr == 0.3
    assert other.states_synced == {0: False, 'k': False}
    both = mxopt.get_updater(mxopt.Adam(learning_rate=0.02))
    both.states = {1: 'x'}
    other.set_states(both.get_states(dump_optimizer=True))
    assert other.states == {1: 'x'} and isinstance(other.optimizer, mxopt.Adam)
    assert other.optimizer.lr == 0.02
    # non-array states are passed through by the context sync
    assert other.sync_state_context(('a', None, 3), None) == ('a', None, 3)
    assert other.sync_state_context(['a'], None) == ['a']


def test_updater_accepts_tuples_and_decodes_bytes_in_place(mx):
    opt = _recording_optimizer(mx, 4)
    upd = mxopt.get_updater(opt)
    upd((0, 1), (_Arr('g0'), _Arr('g1')), (_Arr('w0'), _Arr('w1')))
    assert opt.calls[-1][0] == [0, 1]
    keys = [b'a', 'b']
    upd(keys, [_Arr('ga'), _Arr('gb')], [_Arr('wa'), _Arr('wb')])
    assert keys == ['a', 'b'] and opt.calls[-1][0] == ['a', 'b']
