"""Hierarchical parameter tree: Prior, Param, Node.

Host-side mirror of /root/reference/model.py (same names, argument meaning and error
behaviour), written for this package: priors are closed forms instead of scipy objects and
the tree can be flattened once (flatten.py) so that the per-walker traversal of
Node.__set_parameter_vector__ / ln_prior / ln_like (model.py:382-498,586-603) never runs on
the sampling path.  The scalar methods below remain for set-up, reporting and plotting.
"""
import math
import os
import warnings

import numpy as np

TINY = -np.inf
PRIOR_TYPES = ('gauss', 'gaussPos', 'uniform', 'log_uniform', 'mod_jeff')
_LN_SQRT_2PI = 0.5 * math.log(2.0 * math.pi)
_LN_MIN_DENORMAL = math.log(5e-324)


def extract_par_and_key(key):
    """'wdFlux_long_label' -> ('wdFlux', 'long_label'); the GP hyper-parameters keep their
    three-word names: 'ln_ampin_gp_core' -> ('ln_ampin_gp', 'core') (model.py:22-37)."""
    words = key.split('_')
    n = 3 if key.startswith('ln_') else 1
    return '_'.join(words[:n]), '_'.join(words[n:])


class Prior(object):
    """Prior on one parameter: gauss / gaussPos (mean, sigma), uniform / log_uniform (lo, hi),
    mod_jeff (knee p0, upper limit) -- model.py:40-113."""

    def __init__(self, type, p1, p2):
        assert type in PRIOR_TYPES
        self.type = type
        self.p1 = p1
        self.p2 = p2
        if type == 'log_uniform' and self.p1 < 1.0e-30:
            warnings.warn('lower limit on log_uniform prior rescaled from %f to 1.0e-30' % self.p1)
            self.p1 = 1.0e-30
        self.normalise = 1.0
        if type == 'log_uniform':
            # The reference normalises with |integral of ln_prob| over (p1, p2), taken while
            # normalise is still 1 (model.py:77-79): that is |[x - x ln x]| and not ln(p2/p1).
            # Kept as is: it shifts ln_prior by a constant the chains were produced with.
            prim = lambda x: x - x * math.log(x)
            self.normalise = abs(prim(self.p2) - prim(self.p1))
        elif type == 'mod_jeff':
            self.normalise = math.log((self.p1 + self.p2) / self.p1)

    def ln_prob(self, val):
        kind = self.type
        if kind in ('gauss', 'gaussPos'):
            if kind == 'gaussPos' and val <= 0.0:
                return TINY
            z = (val - self.p1) / self.p2
            t = -0.5 * z * z - _LN_SQRT_2PI
            # scipy's pdf underflows to exactly 0 far in the tails -> -inf (model.py:85-89)
            if not t >= _LN_MIN_DENORMAL:
                return TINY
            return t - math.log(self.p2)
        if kind == 'uniform':
            if self.p1 < val < self.p2:
                return math.log(1.0 / abs(self.p1 - self.p2))
            return TINY
        if kind == 'log_uniform':
            if self.p1 < val < self.p2:
                return math.log(1.0 / self.normalise / val)
            return TINY
        if 0 < val < self.p2:  # mod_jeff
            return math.log(1.0 / self.normalise / (val + self.p1))
        return TINY


class Param(object):
    """A starting value, a current value, a prior and a flag saying whether it varies."""

    def __init__(self, name, startVal, prior, isVar=True):
        self.name = name
        self.startVal = startVal
        self.prior = prior
        self.currVal = startVal
        self.isVar = isVar

    @classmethod
    def fromString(cls, name, parString):
        """'value priorType p1 p2 [isVar]' (model.py:126-137)."""
        f = parString.split()
        isVar = bool(int(f[4])) if len(f) == 5 else True
        return cls(name, float(f[0]), Prior(f[1].strip(), float(f[2]), float(f[3])), isVar)

    @property
    def isValid(self):
        return bool(np.isfinite(self.prior.ln_prob(self.currVal)))


class Node:
    """A node of the model tree: any number of children, at most one parent.  Leaves inherit
    the Params of their ancestors; parameter vectors are read and set from any level in
    depth-first order, a node's own variable Params first (model.py:144-844)."""

    node_par_names = ()

    def __init__(self, label, parameter_objects, parent=None, children=None, DEBUG=None):
        self.children = [] if children is None else children
        self.parent = parent
        if isinstance(DEBUG, bool):
            self.DEBUG = DEBUG
        elif self.parent is not None:
            self.DEBUG = self.parent.DEBUG
        else:
            self.DEBUG = False
        try:
            parameter_objects = list(parameter_objects)
        except TypeError:
            parameter_objects = [parameter_objects]
        if not isinstance(label, str):
            raise TypeError("Label must be a string, not {}".format(type(label)))
        self.label = label
        if len(self.node_par_names) != len(parameter_objects):
            raise TypeError('I recieved the wrong number of parameters! Expect: \n{}\nGot:\n{}'.format(
                self.node_par_names, [getattr(p, 'name') for p in parameter_objects]))
        for par in parameter_objects:
            setattr(self, par.name, par)
        self.log('base.__init__', "Successfully did the base Node init")

    # ---- searching -----------------------------------------------------------------
    def search_par(self, label, name):
        """The Param `name` of the node labelled `label` at or below this node, else None."""
        if self.label == label:
            return getattr(self, name)
        for child in self.children:
            found = child.search_par(label, name)
            if found is not None:
                return found
        return None

    def search_Node(self, class_type, label):
        """The node named '<class_type>_<label>' at or below this node, else None."""
        if self.name == "{}_{}".format(class_type, label):
            return self
        for child in self.children:
            found = child.search_Node(class_type, label)
            if found is not None:
                return found
        return None

    def search_node_type(self, class_type, nodes=None):
        """Set of nodes at or below this one whose class name contains class_type."""
        nodes = set() if nodes is None else nodes
        for child in self.children:
            nodes = nodes.union(child.search_node_type(class_type, nodes))
        if class_type in type(self).__name__:
            nodes.add(self)
        return nodes

    def add_child(self, children):
        if not isinstance(children, list):
            children = [children]
        self.children.extend(children)

    # ---- evaluation ----------------------------------------------------------------
    def __call_recursive_func__(self, name, *args, **kwargs):
        """Sum `name` over the children, stopping at the first infinite running total
        (model.py:382-415)."""
        if self.is_leaf:
            raise NotImplementedError('must overwrite {} on leaf nodes of model'.format(name))
        val = 0.0
        for child in self.children:
            val += getattr(child, name)(*args, **kwargs)
            if np.any(np.isinf(val)):
                return val
        return val

    def chisq(self, *args, **kwargs):
        return self.__call_recursive_func__('chisq', *args, **kwargs)

    def ln_like(self, *args, **kwargs):
        return self.__call_recursive_func__('ln_like', *args, **kwargs)

    def ln_prior(self, verbose=False):
        """Sum of the variable Params' prior log-densities at and below this node; -inf as soon
        as any Param (variable or not) violates its prior (model.py:426-474)."""
        lnp = 0.0
        for par in (getattr(self, n) for n in self.node_par_names):
            lp = par.prior.ln_prob(par.currVal)
            if not np.isfinite(lp):
                if verbose:
                    print("Param {} in {} is invalid!".format(par.name, self.name))
                return -np.inf
            if par.isVar:
                lnp += lp
        if verbose:
            print("{} has the following Params:".format(self.name))
            for i in range(0, len(self.node_par_names), 4):
                print(self.node_par_names[i:i + 4])
            print("The sum of parameter ln_priors of {} is {:.3f}\n".format(self.name, lnp))
        for child in self.children:
            lnp += child.ln_prior(verbose=verbose)
            if np.isinf(lnp):
                return lnp
        return lnp

    def ln_prob(self, verbose=False):
        """ln_prior, plus ln_like when the prior is finite; any failure of ln_like -> -inf
        (model.py:476-498)."""
        lnp = self.ln_prior(verbose=verbose)
        if not np.isfinite(lnp):
            if verbose:
                print("{} ln_prior returned infinite!".format(self.name))
            return lnp
        try:
            return lnp + self.ln_like()
        except Exception:
            if verbose:
                print("Failed to evaluate ln_like at {}".format(self.name))
            return -np.inf

    # ---- parameter bookkeeping -------------------------------------------------------
    def __get_inherited_parameter_names__(self):
        names = list(self.node_par_names)
        if self.parent is not None:
            names += self.parent.__get_inherited_parameter_names__()
        return names

    def __get_inherited_parameter_vector__(self):
        vector = [getattr(self, n) for n in self.node_par_names]
        if self.parent is not None:
            vector += self.parent.__get_inherited_parameter_vector__()
        return vector

    def __get_descendant_params__(self):
        """(Params, owning node labels) at and below this node, depth first."""
        params = [getattr(self, n) for n in self.node_par_names]
        labels = [self.label] * len(params)
        for child in self.children:
            p, l = child.__get_descendant_params__()
            params.extend(p)
            labels.extend(l)
        return params, labels

    def __get_descendant_parameter_vector__(self):
        return [p.currVal for p in self.__get_descendant_params__()[0] if p.isVar]

    def __get_descendant_parameter_names__(self):
        params, labels = self.__get_descendant_params__()
        return [p.name + "_" + l for p, l in zip(params, labels) if p.isVar]

    def __set_parameter_vector__(self, vector_values):
        """Consume values from the back: last child first, this node's own Params last
        (model.py:586-603).  Returns what is left for the nodes before."""
        vector = list(vector_values)
        for child in reversed(self.children):
            vector = child.__set_parameter_vector__(vector)
        own = self.node_varpars
        for name, val in zip(reversed(own), reversed(vector)):
            getattr(self, name).currVal = val
        return vector[:len(vector) - len(own)]

    def __check_par_assignments__(self):
        for key in self.node_par_names:
            if key != getattr(self, key).name:
                raise NameError("Incorrect parameter name, {} assigned to {}. \nParameters are taken in the order {}".format(
                    getattr(self, key).name, key, self.node_par_names))

    def __getitem__(self, index):
        name, label = extract_par_and_key(index)
        return self.search_par(label, name)

    def __setitem__(self, index, value):
        name, label = extract_par_and_key(index)
        self.search_par(label, name).currVal = value

    # ---- properties ----------------------------------------------------------------
    @property
    def name(self):
        return "{}_{}".format(type(self).__name__, self.label)

    @property
    def parent(self):
        return self.__parent

    @parent.setter
    def parent(self, parent):
        self.__parent = parent
        if parent is not None:
            parent.add_child(self)

    @property
    def children(self):
        return self.__children

    @children.setter
    def children(self, children):
        if not isinstance(children, list):
            children = list(children)
        self.__children = children
        for child in children:
            child.__parent = self

    @property
    def dynasty_par_names(self):
        return self.__get_descendant_parameter_names__()

    @property
    def dynasty_par_vals(self):
        return self.__get_descendant_parameter_vector__()

    @dynasty_par_vals.setter
    def dynasty_par_vals(self, dynasty_par_vals):
        expect = len(self.dynasty_par_vals)
        if len(dynasty_par_vals) != expect:
            raise ValueError('Wrong vector length on {} - Expected {}, got {}'.format(
                self.name, expect, len(dynasty_par_vals)))
        self.__set_parameter_vector__(dynasty_par_vals)

    @property
    def dynasty_par_dict(self):
        return dict(zip(self.dynasty_par_names, self.dynasty_par_vals))

    @dynasty_par_dict.setter
    def dynasty_par_dict(self, par_dict):
        for key, value in par_dict.items():
            try:
                self[key].currVal = value
            except AttributeError as e:
                print(repr(e))

    @property
    def ancestor_param_dict(self):
        """{name: Param} of this node and everything above it, variable or not."""
        return dict(zip(self.__get_inherited_parameter_names__(), self.__get_inherited_parameter_vector__()))

    @property
    def ancestor_par_names(self):
        return self.__get_inherited_parameter_names__()

    @property
    def node_varpars(self):
        return [n for n in self.node_par_names if getattr(self, n).isVar]

    @property
    def is_root(self):
        return self.parent is None

    @property
    def is_leaf(self):
        return len(self.children) == 0

    # ---- diagnostics ---------------------------------------------------------------
    @property
    def structure(self):
        """Nested dict of the tree below this node (networkx tree_data layout)."""
        return {"id": self.name, "children": [c.structure for c in self.children]} if self.children else {"id": self.name}

    @property
    def DEBUG(self):
        return self.__DEBUG

    @DEBUG.setter
    def DEBUG(self, flag):
        for child in self.children:
            child.DEBUG = flag
        self.__DEBUG = flag

    def log(self, called_by, message='\n', log_stack=False):
        """Append to DEBUGGING/<pid>.txt when DEBUG is on (model.py:763-793).  Callers pass
        plain strings; nothing is formatted when DEBUG is off."""
        if not self.DEBUG:
            return
        os.makedirs("DEBUGGING", exist_ok=True)
        if callable(message):
            message = message()
        if not message.endswith('\n'):
            message += "\n"
        with open(os.path.join('DEBUGGING', "{}.txt".format(os.getpid())), 'a+') as f:
            f.write('*' * 150 + "\n")
            f.write("--> Logger called by function {} in node {}\n".format(called_by, self.name))
            if log_stack:
                import inspect
                stack = "\n     ".join("File {}, line {}, function {}".format(x.filename, x.lineno, x.function)
                                       for x in reversed(inspect.stack()))
                f.write("--> The function stack is \n     {}\n".format(stack))
            f.write(message)
            f.write('~' * 150 + "\n\n\n")

    def report_relatives(self):
        print("Reporting family tree of {}:".format(self.name))
        print("    Parent: {}".format(self.parent.name if self.parent is not None else 'None'))
        print("    Children:")
        for child in self.children:
            print("      {}".format(child.name))
            for grandchild in child.children:
                print("       - {}".format(grandchild.name))

    def report(self, also_relatives=True):
        if also_relatives:
            self.report_relatives()
        print("  Parameter vector, and labels:")
        for par, val in zip(self.dynasty_par_names, self.dynasty_par_vals):
            print("  {:>10s} = {:<.3f}".format(par, val))
        print("\n")

    def create_tree(self, G=None, called=True):
        """networkx DiGraph of the tree below this node (model.py:819-844)."""
        import networkx as nx
        if called:
            G = nx.DiGraph()
        G.add_node(self.name)
        for child in self.children:
            child.create_tree(G, called=False)
            G.add_edge(self.name, child.name)
        if called:
            self.nx_graph = G
        return G
