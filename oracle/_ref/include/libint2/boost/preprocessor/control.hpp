# /* **************************************************************************
#  *                                                                          *
#  *     (C) Copyright Paul Mensonides 2002.
#  *     Distributed under the Boost Software License, Version 1.0. (See
#  *     accompanying file LICENSE_1_0.txt or copy at
#  *     http://www.boost.org/LICENSE_1_0.txt)
#  *                                                                          *
#  ************************************************************************** */
#
# /* See http://www.boost.org for most recent version. */
#
# ifndef BOOST_PREPROCESSOR_CONTROL_HPP
# define BOOST_PREPROCESSOR_CONTROL_HPP
#
# include <libint2/boost/preprocessor/control/deduce_d.hpp>
# include <libint2/boost/preprocessor/control/expr_if.hpp>
# include <libint2/boost/preprocessor/control/expr_iif.hpp>
# include <libint2/boost/preprocessor/control/if.hpp>
# include <libint2/boost/preprocessor/control/iif.hpp>
# include <libint2/boost/preprocessor/control/while.hpp>
#
# endif
