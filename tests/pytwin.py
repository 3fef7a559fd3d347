"""Slow pure-Python twin of the reference's search()/simulate() (src/solver.jl:41-157) for the two
Bool-state problems, written independently of oracle/pomcp_oracle.c (recursive, dict-based like the
Julia code) but on the same Philox draw map.  Used only to cross-check the C oracle on tiny cases."""
import math

from pomcp_jl_b200.problems import philox4x32_10

TAG_ROOT, TAG_TREE, TAG_ROLLOUT = 0, 1, 2
U = 2.0 ** -32


class Node:
    def __init__(self):
        self.N = 0
        self.children = {}  # action -> ActNode (insertion = ascending action order)
        self.particles = []


class ActNode:
    def __init__(self, a, N, V):
        self.a, self.N, self.V, self.children = a, N, V, {}


class Twin:
    def __init__(self, pomdp, c, seed, eps=0.01, max_depth=2 ** 62, log_fn=math.log, fwc=False, fo_fwc=False):
        self.p, self.c, self.eps, self.max_depth = pomdp, c, eps, max_depth
        self.key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
        self.gamma = pomdp.discount
        self.root = Node()
        self.epoch = 0
        self.log = log_fn
        self.kind = type(pomdp).__name__
        self.fwc = fwc  # Baby: RolloutEstimator(FeedWhenCrying()), test/runtests.jl:13-19
        self.fo_fwc = fo_fwc  # Baby: FORollout(FeedWhenCrying()), test/runtests.jl:25-31

    # ---- generative model (Tiger / Baby), transition draw first, observation draw second
    def step(self, s, a, ut, uo, need_obs=True):
        p = self.p
        if self.kind == "TigerPOMDP":
            if a == 0:
                return s, (s if uo < p.p_listen_correctly else 1 - s), p.r_listen
            opened_tiger = s if a == 1 else 1 - s
            r = p.r_findtiger if opened_tiger else p.r_escapetiger
            return (1 if ut < 0.5 else 0), (1 if uo < 0.5 else 0), r
        r = (p.r_hungry if s else 0.0) + (p.r_feed if a else 0.0)
        sp = 0 if a else (1 if s else (1 if ut < p.p_become_hungry else 0))
        pc = p.p_cry_when_hungry if sp else p.p_cry_when_not_hungry
        return sp, (1 if uo < pc else 0), r

    def initial(self, u0):
        return (1 if u0 < 0.5 else 0) if self.kind == "TigerPOMDP" else 0

    def rollout(self, s, steps, q):
        disc, total, step, blk, w = 1.0, 0.0, 0, -1, None
        nA = self.p.n_actions
        while step < steps:
            word = step * 2
            if word >> 2 != blk:
                blk = word >> 2
                w = philox4x32_10((blk, TAG_ROLLOUT, q, self.epoch), self.key)
            a = (w[word & 3] * nA) >> 32
            sp, _, r = self.step(s, a, w[(word & 3) + 1] * U, 0.0)
            total += disc * r
            s = sp
            disc *= self.gamma
            step += 1
        return total

    def rollout_fwc(self, s, steps, q, last_obs):
        """feed iff the previous observation was crying; starts from the leaf's observation label
        (src/rollout.jl:105-106).  Two words per step: transition, observation."""
        disc, total, step, blk, w = 1.0, 0.0, 0, -1, None
        while step < steps:
            word = step * 2
            if word >> 2 != blk:
                blk = word >> 2
                w = philox4x32_10((blk, TAG_ROLLOUT, q, self.epoch), self.key)
            a = 1 if last_obs else 0
            sp, o, r = self.step(s, a, w[word & 3] * U, w[(word & 3) + 1] * U)
            total += disc * r
            s, last_obs = sp, o
            disc *= self.gamma
            step += 1
        return total

    def rollout_fo_fwc(self, s, steps, q):
        """MDP rollout, the policy is shown the state: feed iff hungry.  One word per step (transition)."""
        disc, total, step, blk, w = 1.0, 0.0, 0, -1, None
        while step < steps:
            if step >> 2 != blk:
                blk = step >> 2
                w = philox4x32_10((blk, TAG_ROLLOUT, q, self.epoch), self.key)
            sp, _, r = self.step(s, 1 if s else 0, w[step & 3] * U, 0.0)
            total += disc * r
            s = sp
            disc *= self.gamma
            step += 1
        return total

    def simulate(self, h, s, depth, q, label=0):
        if self.gamma ** depth < self.eps or depth >= self.max_depth:
            return 0.0
        if not h.children:
            for a in range(self.p.n_actions):
                h.children[a] = ActNode(a, 0, 0.0)
            if depth > 0:
                steps = min(self.max_depth - depth, math.ceil(math.log(self.eps) / math.log(self.gamma) - depth))
                if self.fwc:
                    return self.gamma ** depth * self.rollout_fwc(s, steps, q, label)
                if self.fo_fwc:
                    return self.gamma ** depth * self.rollout_fo_fwc(s, steps, q)
                return self.gamma ** depth * self.rollout(s, steps, q)
            return 0.0
        best, best_node = -math.inf, None
        for node in h.children.values():
            if node.N == 0 and h.N == 1:
                crit = node.V
            elif node.N == 0 and node.V == -math.inf:
                crit = math.inf
            elif node.N == 0:
                crit = math.inf if self.c > 0 else math.nan
            else:
                crit = node.V + self.c * math.sqrt(self.log(h.N) / node.N)
            if crit >= best:
                best, best_node = crit, node
        w = philox4x32_10((depth, TAG_TREE, q, self.epoch), self.key)
        sp, o, r = self.step(s, best_node.a, w[0] * U, w[2] * U)
        if o not in best_node.children:
            best_node.children[o] = Node()
        hao = best_node.children[o]
        R = r + self.gamma * self.simulate(hao, sp, depth + 1, q, o)
        hao.particles.append(sp)
        hao.N += 1
        best_node.N += 1
        best_node.V += (R - best_node.V) / best_node.N
        return R

    def search(self, Q):
        for q in range(Q):
            w = philox4x32_10((0, TAG_ROOT, q, self.epoch), self.key)
            s = self.initial(w[0] * U)
            self.simulate(self.root, s, 0, q)
            self.root.N += 1
        self.epoch += 1
        nodes = list(self.root.children.values())
        best = nodes[0]
        for n in nodes[1:]:
            if n.V >= best.V:
                best = n
        return best.a, [n.N for n in nodes], [n.V for n in nodes]


def rocksample_rollout(pomdp, state, steps, key, q, tagword=TAG_ROLLOUT, epoch=0):
    """Random-policy rollout of RockSample written from the problem statement (Smith & Simmons 2004; moves
    0 north y+1, 1 east x+1 (off the right edge = exit, terminal), 2 south, 3 west, 4 sample, 5+i check rock i),
    on the rollout draw map of DESIGN.md §2: block = step // 8 of the (block, tag, query, epoch) Philox
    stream, 16-bit uniform = low half of word (step % 8) // 2 for even steps, high half for odd ones,
    a = floor(u16 * nA / 2^16).  Returns (discounted return, steps taken)."""
    n, nA = pomdp.n, pomdp.n_actions
    x, y, rocks = state & 15, (state >> 4) & 15, (state >> 8) & 0xFFFF
    if state & 0x80000000:
        return 0.0, 0
    rock_at = {xy: i for i, xy in enumerate(pomdp.rocks)}
    disc, total, step, blk, w = 1.0, 0.0, 0, -1, None
    while disc > 0.0 and step < steps:
        if step // 8 != blk:
            blk = step // 8
            w = philox4x32_10((blk, tagword, q, epoch), key)
        word = w[(step % 8) // 2]
        u16 = (word >> 16) if (step & 1) else (word & 0xFFFF)
        a = (u16 * nA) >> 16
        r, done = 0.0, False
        if a == 0:
            y = min(y + 1, n - 1)
        elif a == 1:
            if x + 1 > n - 1:
                r, done = pomdp.r_exit, True
            else:
                x += 1
        elif a == 2:
            y = max(y - 1, 0)
        elif a == 3:
            x = max(x - 1, 0)
        elif a == 4 and (x, y) in rock_at:
            i = rock_at[(x, y)]
            r = pomdp.r_good if (rocks >> i) & 1 else pomdp.r_bad
            rocks &= ~(1 << i)
        total += disc * r
        disc *= pomdp.discount
        step += 1
        if done:
            break
    return total, step
