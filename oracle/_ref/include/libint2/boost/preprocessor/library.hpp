# /* **************************************************************************
#  *                                                                          *
#  *     (C) Copyright Paul Mensonides 2002-2011.                             *
#  *     (C) Copyright Edward Diener 2011.                                    *
#  *     Distributed under the Boost Software License, Version 1.0. (See      *
#  *     accompanying file LICENSE_1_0.txt or copy at                         *
#  *     http://www.boost.org/LICENSE_1_0.txt)                                *
#  *                                                                          *
#  ************************************************************************** */
#
# /* See http://www.boost.org for most recent version. */
#
# ifndef BOOST_PREPROCESSOR_LIBRARY_HPP
# define BOOST_PREPROCESSOR_LIBRARY_HPP
#
# include <libint2/boost/preprocessor/arithmetic.hpp>
# include <libint2/boost/preprocessor/array.hpp>
# include <libint2/boost/preprocessor/cat.hpp>
# include <libint2/boost/preprocessor/comparison.hpp>
# include <libint2/boost/preprocessor/config/limits.hpp>
# include <libint2/boost/preprocessor/control.hpp>
# include <libint2/boost/preprocessor/debug.hpp>
# include <libint2/boost/preprocessor/facilities.hpp>
# include <libint2/boost/preprocessor/iteration.hpp>
# include <libint2/boost/preprocessor/list.hpp>
# include <libint2/boost/preprocessor/logical.hpp>
# include <libint2/boost/preprocessor/punctuation.hpp>
# include <libint2/boost/preprocessor/repetition.hpp>
# include <libint2/boost/preprocessor/selection.hpp>
# include <libint2/boost/preprocessor/seq.hpp>
# include <libint2/boost/preprocessor/slot.hpp>
# include <libint2/boost/preprocessor/stringize.hpp>
# include <libint2/boost/preprocessor/tuple.hpp>
# include <libint2/boost/preprocessor/variadic.hpp>
# include <libint2/boost/preprocessor/wstringize.hpp>
#
# endif
